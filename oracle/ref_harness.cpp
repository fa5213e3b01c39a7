// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Links the UNMODIFIED reference (oracle/_ref/libpbrt_ref.a, built by
// oracle/Makefile from /root/reference) and drives its own classes to produce
// golden vectors for the volumetric photon-mapping path:
//   * KdTree<Photon>::Lookup + PhotonProcess      (core/kdtree.h:150-183,
//                                                  core/photonshooter.h:186-203)
//   * PhotonVolumeIntegrator::LPhoton / Li / Transmittance
//                                                 (integrators/photonvolume.cpp)
//   * Scene::Intersect / IntersectP               (accelerators/bvh.cpp:585-685)
//   * PhotonShootingTask::Run                     (core/photonshooter.cpp:232-357)
//   * SingleScatteringIntegrator::Li / EmissionIntegrator::Li
//                                                 (integrators/single.cpp:66-138, emission.cpp:63-106)
//   * the flattened scene the C ABI consumes      
//
// It reaches private members with `#define private public` and reaches the
// file-static RenderOptions by #including the reference's core/api.cpp as part
// of this translation unit (which therefore replaces core_api.o at link time).
// The parser's call to pbrtWorldEnd() lands in the harness version below.
// No reference source is copied or modified.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <vector>
#include <string>
#include <map>
#include <set>
#include <list>
#include <algorithm>
#include <sstream>
#include <iostream>
#include <fstream>
#include <memory>
#include <typeinfo>
#include <sys/time.h>

#define private public
#define protected public
#define pbrtWorldEnd pbrtWorldEnd_reference
#include "core/api.cpp"
#undef pbrtWorldEnd
#include "integrators/photonmap.cpp"
#include "core/photonshooter.h"
#include "core/kdtree.h"
#include "core/parser.h"
#include "core/scene.h"
#include "core/light.h"
#include "core/sampler.h"
#include "core/camera.h"
#include "core/intersection.h"
#include "accelerators/bvh.h"
#include "accelerators/kdtreeaccel.h"
#include "accelerators/grid.h"
#include "shapes/trianglemesh.h"
#include "shapes/sphere.h"
#include "lights/point.h"
#include "lights/spot.h"
#include "lights/distant.h"
#include "lights/diffuse.h"
#include "volumes/homogeneous.h"
#include "volumes/volumegrid.h"
#include "volumes/exponential.h"
#include "volumes/rainbow.h"
#include "materials/matte.h"
#include "materials/glass.h"
#include "integrators/photonvolume.h"
#include "renderers/samplerrenderer.h"
#undef private
#undef protected

#include "../include/pv.h"

#include "../cs348b-pbrt_b200/host/pv_export.inl"

static std::vector<std::string> g_ops;
static int g_rc = 0;

static double now_s() {
    struct timeval tv; gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

template <typename T> static void wr(FILE *f, const T *p, size_t n) {
    if (n && fwrite(p, sizeof(T), n, f) != n) { perror("fwrite"); exit(3); }
}
template <typename T> static void rd(FILE *f, T *p, size_t n) {
    if (n && fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(3); }
}
static FILE *xopen(const std::string &fn, const char *mode) {
    FILE *f = fopen(fn.c_str(), mode);
    if (!f) { perror(fn.c_str()); exit(3); }
    return f;
}
static uint64_t read_header(FILE *f, const char *magic) {
    char m[8]; uint64_t n;
    rd(f, m, 8); rd(f, &n, 1);
    if (memcmp(m, magic, 8)) { fprintf(stderr, "bad magic, want %s\n", magic); exit(3); }
    return n;
}
static void write_header(FILE *f, const char *magic, uint64_t n) {
    wr(f, magic, 8); wr(f, &n, 1);
}
// ---------------------------------------------------------------- photons
static std::vector<Photon> g_photons;       // original (merge) order
static std::vector<Photon> g_surf[3];       // caustic, indirect, direct photons of the last --shoot, merge order
static std::vector<RadiancePhoton> g_rad;   // radiance photons with Lo (ComputeRadianceTask)
static std::vector<Spectrum> g_rho_r, g_rho_t;
static int g_paths[4] = {0, 0, 0, 0};       // nCausticPaths, nIndirectPaths, nDirectPaths, nVolumePaths
static KdTree<RadiancePhoton> *g_radMap = NULL;   // the reference's radiance map, photons tagged with their list index
// PhotonIntegrator's LPhoton (integrators/photonmap.cpp:62) is file-static: the reference's photonmap.cpp is compiled as part of
// this translation unit (see the #include next to core/api.cpp), which makes it callable here; the archive's copy of that
// object is then never pulled in by the linker.
static uint32_t g_nshot = 0;
static double g_shoot_seconds = 0;

static void stash_indices(std::vector<Photon> &v) {
    // The fork's Spectrum carries an unused `intensity` float; park the
    // original index there (bit pattern) so kd-tree node order can be mapped back.
    for (uint32_t i = 0; i < v.size(); ++i) memcpy(&v[i].alpha.intensity, &i, 4);
}
static uint32_t photon_index(const Photon *p) {
    uint32_t i; memcpy(&i, &p->alpha.intensity, 4); return i;
}
static void install_volume_map(PhotonShooter *sh) {
    stash_indices(g_photons);
    delete sh->volumeMap; sh->volumeMap = NULL;
    if (g_photons.size()) sh->volumeMap = new KdTree<Photon>(g_photons);
}

// Restates only the DRIVER part of PhotonShooter::Preprocess
// (core/photonshooter.cpp:457-503) so that nshot and the pre-kd-tree photon
// order are observable; the tasks themselves are the reference's own.
static void shoot(PhotonShooter *sh, const Scene *scene, const Camera *camera, const Renderer *renderer) {
    if (scene->lights.size() == 0) return;
    Mutex *mutex = Mutex::Create();
    int nDirectPaths = 0;
    vector<Photon> causticPhotons, directPhotons, indirectPhotons, volumePhotons;
    vector<RadiancePhoton> radiancePhotons;
    vector<Spectrum> rpReflectances, rpTransmittances;
    bool abortTasks = false;
    uint32_t nshot = 0;
    Distribution1D *lightDistribution = ComputeLightSamplingCDF(scene);
    ProgressReporter progress(sh->nCausticPhotonsWanted + sh->nIndirectPhotonsWanted + sh->nVolumePhotonsWanted, "Shooting photons");
    vector<Task *> tasks;
    int nTasks = NumSystemCores();
    for (int i = 0; i < nTasks; ++i)
        tasks.push_back(new PhotonShootingTask(i, camera ? camera->shutterOpen : 0.f, *mutex, sh, progress,
            abortTasks, nDirectPaths, directPhotons, indirectPhotons, causticPhotons, volumePhotons,
            radiancePhotons, rpReflectances, rpTransmittances, nshot, lightDistribution, scene, renderer));
    double t0 = now_s();
    EnqueueTasks(tasks);
    WaitForAllTasks();
    g_shoot_seconds = now_s() - t0;
    for (size_t i = 0; i < tasks.size(); ++i) delete tasks[i];
    Mutex::Destroy(mutex);
    progress.Done();
    if (causticPhotons.size()) sh->causticMap = new KdTree<Photon>(causticPhotons);
    if (indirectPhotons.size()) sh->indirectMap = new KdTree<Photon>(indirectPhotons);
    // radiance photons exactly as PhotonShooter::Preprocess computes them (core/photonshooter.cpp:494-524)
    KdTree<Photon> *directMap = directPhotons.size() ? new KdTree<Photon>(directPhotons) : NULL;
    if (sh->finalGather && radiancePhotons.size()) {
        vector<Task *> radianceTasks;
        uint32_t numTasks = 64;
        ProgressReporter progRadiance(numTasks, "Computing photon radiances");
        for (uint32_t i = 0; i < numTasks; ++i)
            radianceTasks.push_back(new ComputeRadianceTask(progRadiance, i, numTasks, radiancePhotons, rpReflectances, rpTransmittances,
                sh->nLookup, sh->maxDistSquared, nDirectPaths, directMap, sh->nIndirectPaths, sh->indirectMap,
                sh->nCausticPaths, sh->causticMap, sh->nVolumePaths, sh->volumeMap));
        EnqueueTasks(radianceTasks);
        WaitForAllTasks();
        for (uint32_t i = 0; i < radianceTasks.size(); ++i) delete radianceTasks[i];
        progRadiance.Done();
    }
    delete directMap;
    g_surf[0] = causticPhotons; g_surf[1] = indirectPhotons; g_surf[2] = directPhotons;
    g_rad = radiancePhotons; g_rho_r = rpReflectances; g_rho_t = rpTransmittances;
    delete g_radMap; g_radMap = NULL;
    if (radiancePhotons.size()) {
        vector<RadiancePhoton> tagged = radiancePhotons;
        for (uint32_t i = 0; i < tagged.size(); ++i) memcpy(&tagged[i].Lo.intensity, &i, 4);     // unused member of the fork's Spectrum
        g_radMap = new KdTree<RadiancePhoton>(tagged);
    }
    g_paths[0] = sh->nCausticPaths; g_paths[1] = sh->nIndirectPaths; g_paths[2] = nDirectPaths; g_paths[3] = sh->nVolumePaths;
    g_photons.swap(volumePhotons);
    g_nshot = nshot;
    install_volume_map(sh);
    fprintf(stderr, "[harness] shot: nshot=%u volume=%zu caustic=%zu indirect=%zu tasks=%d seconds=%.3f\n",
            nshot, g_photons.size(), causticPhotons.size(), indirectPhotons.size(), nTasks, g_shoot_seconds);
}

static void load_photons(const std::string &fn) {
    FILE *f = xopen(fn, "rb");
    uint64_t n = read_header(f, "PVPHOT01");
    g_photons.resize(n);
    std::vector<float> rec(36);
    for (uint64_t i = 0; i < n; ++i) {
        rd(f, rec.data(), 36);
        Photon &p = g_photons[i];
        p.p = Point(rec[0], rec[1], rec[2]);
        p.wi = Vector(rec[3], rec[4], rec[5]);
        p.alpha = Spectrum(0.f);
        memcpy(p.alpha.c, &rec[6], 30 * sizeof(float));
        p.alpha.lambda = -1.f;
    }
    fclose(f);
}
static void dump_photon_vec(const std::string &fn, const std::vector<Photon> &v);
static void dump_photons(const std::string &fn) { dump_photon_vec(fn, g_photons); }
// surface maps of the last --shoot: <prefix>.caustic / .indirect / .direct (PVPHOT01) and <prefix>.radiance
// (PVRADP01: p[3] n[3] Lo[30] rho_r[30] rho_t[30] per radiance photon)
static void dump_maps(const std::string &prefix) {
    static const char *names[3] = {".caustic", ".indirect", ".direct"};
    for (int k = 0; k < 3; ++k) dump_photon_vec(prefix + names[k], g_surf[k]);
    FILE *f = xopen(prefix + ".radiance", "wb");
    write_header(f, "PVRADP01", g_rad.size());
    for (size_t i = 0; i < g_rad.size(); ++i) {
        float rec[96];
        const RadiancePhoton &r = g_rad[i];
        rec[0] = r.p.x; rec[1] = r.p.y; rec[2] = r.p.z; rec[3] = r.n.x; rec[4] = r.n.y; rec[5] = r.n.z;
        memcpy(&rec[6], r.Lo.c, 30 * sizeof(float)); memcpy(&rec[36], g_rho_r[i].c, 30 * sizeof(float)); memcpy(&rec[66], g_rho_t[i].c, 30 * sizeof(float));
        wr(f, rec, 96);
    }
    fclose(f);
}
static void dump_photon_vec(const std::string &fn, const std::vector<Photon> &g_photons) {
    FILE *f = xopen(fn, "wb");
    write_header(f, "PVPHOT01", g_photons.size());
    for (size_t i = 0; i < g_photons.size(); ++i) {
        float rec[36];
        const Photon &p = g_photons[i];
        rec[0] = p.p.x; rec[1] = p.p.y; rec[2] = p.p.z;
        rec[3] = p.wi.x; rec[4] = p.wi.y; rec[5] = p.wi.z;
        memcpy(&rec[6], p.alpha.c, 30 * sizeof(float));
        wr(f, rec, 36);
    }
    fclose(f);
}

struct IdxD2 { uint32_t idx; float d2; };
static bool idxd2_less(const IdxD2 &a, const IdxD2 &b) { return a.d2 == b.d2 ? a.idx < b.idx : a.d2 < b.d2; }

static void knn(PhotonShooter *sh, const std::string &qfn, uint32_t k, float r2, const std::string &ofn) {
    FILE *f = xopen(qfn, "rb");
    uint64_t n = read_header(f, "PVQRY001");
    std::vector<float> q(6 * n); rd(f, q.data(), q.size()); fclose(f);
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVKNN001", n);
    wr(o, &k, 1);
    std::vector<ClosePhoton> buf(k);
    std::vector<IdxD2> res;
    std::vector<uint32_t> idx(k); std::vector<float> d2(k);
    double t0 = now_s();
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t nFound = 0;
        res.clear();
        if (sh->volumeMap) {
            PhotonProcess proc(k, buf.data());
            float md2 = r2;
            sh->volumeMap->Lookup(Point(q[6*i], q[6*i+1], q[6*i+2]), proc, md2);
            nFound = proc.nFound;
            for (uint32_t j = 0; j < nFound; ++j) {
                IdxD2 e = { photon_index(buf[j].photon), buf[j].distanceSquared };
                res.push_back(e);
            }
            std::sort(res.begin(), res.end(), idxd2_less);
        }
        for (uint32_t j = 0; j < k; ++j) {
            idx[j] = j < nFound ? res[j].idx : 0xFFFFFFFFu;
            d2[j] = j < nFound ? res[j].d2 : INFINITY;
        }
        wr(o, &nFound, 1); wr(o, idx.data(), k); wr(o, d2.data(), k);
    }
    fprintf(stderr, "[harness] knn: %llu queries k=%u in %.3f s\n", (unsigned long long)n, k, now_s() - t0);
    fclose(o);
}

static void lphoton(PhotonVolumeIntegrator *vi, const Scene *scene, const std::string &qfn, const std::string &ofn) {
    FILE *f = xopen(qfn, "rb");
    uint64_t n = read_header(f, "PVQRY001");
    std::vector<float> q(6 * n); rd(f, q.data(), q.size()); fclose(f);
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVSPEC01", n);
    for (uint64_t i = 0; i < n; ++i) {
        Spectrum L = vi->LPhoton(vi->photonShooter->volumeMap, vi->nUsed, NULL,
                                 Vector(q[6*i+3], q[6*i+4], q[6*i+5]), Point(q[6*i], q[6*i+1], q[6*i+2]),
                                 scene->volumeRegion, vi->maxDistSquared, 0.f);
        wr(o, L.c, 30);
    }
    fclose(o);
}

// The surface integrator's two photon lookups, driven through the reference's own code on the maps of the last --shoot.
// (1) LPhoton (integrators/photonmap.cpp:62-108) at (p, n) with a purely diffuse-reflective and a purely diffuse-transmissive BSDF
//     of unit albedo: the two calls return Lr / pi and Lt / pi of its diffuse branch.
static void surface_lphoton(PhotonShooter *sh, const std::string &which, const std::string &qfn, const std::string &ofn) {
    FILE *f = xopen(qfn, "rb");
    uint64_t n = read_header(f, "PVQRY001");
    std::vector<float> q(6 * n); rd(f, q.data(), q.size()); fclose(f);
    KdTree<Photon> *map = which == "caustic" ? sh->causticMap : sh->indirectMap;
    int nPaths = which == "caustic" ? sh->nCausticPaths : sh->nIndirectPaths;
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVSLPH01", n);
    std::vector<ClosePhoton> buf(sh->nLookup);
    MemoryArena arena;
    RNG rng(7);
    for (uint64_t i = 0; i < n; ++i) {
        Point p(q[6*i], q[6*i+1], q[6*i+2]); Normal nn(q[6*i+3], q[6*i+4], q[6*i+5]);
        Vector v2, v3; CoordinateSystem(Vector(nn), &v2, &v3);
        DifferentialGeometry dg(p, v2, v3, Normal(0, 0, 0), Normal(0, 0, 0), 0.f, 0.f, NULL);
        dg.nn = nn;                                              // exactly the query normal
        Intersection isect; isect.dg = dg;
        Spectrum out[2];
        for (int t = 0; t < 2; ++t) {
            BSDF *bsdf = BSDF_ALLOC(arena, BSDF)(dg, dg.nn);
            BxDF *lam = BSDF_ALLOC(arena, Lambertian)(Spectrum(1.f));
            bsdf->Add(t == 0 ? lam : (BxDF *)BSDF_ALLOC(arena, BRDFToBTDF)(lam));
            out[t] = map ? LPhoton(map, nPaths, sh->nLookup, buf.data(), bsdf, rng, isect, Vector(nn), sh->maxDistSquared) : Spectrum(0.f);
        }
        wr(o, out[0].c, 30); wr(o, out[1].c, 30);
        arena.FreeAll();
    }
    fclose(o);
}
// (2) the radiance-photon lookup of final gathering (integrators/photonmap.cpp:238-243): RadiancePhotonProcess, unbounded radius
static void radiance_nearest(const std::string &qfn, const std::string &ofn) {
    FILE *f = xopen(qfn, "rb");
    uint64_t n = read_header(f, "PVQRY001");
    std::vector<float> q(6 * n); rd(f, q.data(), q.size()); fclose(f);
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVRADN01", n);
    for (uint64_t i = 0; i < n; ++i) {
        Normal nn(q[6*i+3], q[6*i+4], q[6*i+5]);
        RadiancePhotonProcess proc(nn);
        float md2 = INFINITY;
        if (g_radMap) g_radMap->Lookup(Point(q[6*i], q[6*i+1], q[6*i+2]), proc, md2);
        uint32_t idx = 0xFFFFFFFFu;
        Spectrum Lo(0.f);
        if (proc.photon) { memcpy(&idx, &proc.photon->Lo.intensity, 4); Lo = proc.photon->Lo; }
        wr(o, &idx, 1); wr(o, &md2, 1); wr(o, Lo.c, 30);
    }
    fclose(o);
}

static std::vector<pv_ray> read_rays(const std::string &fn) {
    FILE *f = xopen(fn, "rb");
    uint64_t n = read_header(f, "PVRAY001");
    std::vector<pv_ray> r(n); rd(f, r.data(), n); fclose(f);
    return r;
}
static Ray to_ray(const pv_ray &r) {
    return Ray(Point(r.o[0], r.o[1], r.o[2]), Vector(r.d[0], r.d[1], r.d[2]), r.mint, r.maxt, r.time);
}

static void intersect(const Scene *scene, const std::string &rfn, const std::string &ofn) {
    std::vector<pv_ray> rays = read_rays(rfn);
    BVHAccel *bvh = dynamic_cast<BVHAccel *>(scene->aggregate);
    std::map<const Primitive *, uint32_t> index;
    for (uint32_t i = 0; i < bvh->primitives.size(); ++i) index[bvh->primitives[i].GetPtr()] = i;
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVHIT001", rays.size());
    for (size_t i = 0; i < rays.size(); ++i) {
        Ray r = to_ray(rays[i]);
        Intersection isect;
        uint32_t prim = 0xFFFFFFFFu; float t = INFINITY;
        if (scene->Intersect(r, &isect)) { prim = index[isect.primitive]; t = r.maxt; }
        Ray r2 = to_ray(rays[i]);
        uint32_t occl = scene->IntersectP(r2) ? 1 : 0;
        wr(o, &prim, 1); wr(o, &t, 1); wr(o, &occl, 1);
    }
    fclose(o);
}

static void li(SamplerRenderer *ren, const Scene *scene, const std::string &rfn, uint32_t seed, const std::string &ofn) {
    std::vector<pv_ray> rays = read_rays(rfn);
    PhotonVolumeIntegrator *vi = dynamic_cast<PhotonVolumeIntegrator *>(ren->volumeIntegrator);
    if (!vi) { fprintf(stderr, "volume integrator is not photonvolume\n"); exit(4); }
    Sample *sample = new Sample(ren->sampler, ren->surfaceIntegrator, ren->volumeIntegrator, scene);
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVLI0001", rays.size());
    MemoryArena arena;
    double t0 = now_s();
    for (size_t i = 0; i < rays.size(); ++i) {
        RayDifferential r(to_ray(rays[i]));
        RNG rng(seed + (uint32_t)i);
        sample->oneD[vi->scatterSampleOffset][0] = rays[i].u_scatter;
        sample->oneD[vi->tauSampleOffset][0] = 0.5f;
        Spectrum T(1.f);
        Spectrum L = vi->Li(scene, ren, r, sample, rng, &T, arena);
        wr(o, L.c, 30); wr(o, T.c, 30);
        arena.FreeAll();
    }
    fprintf(stderr, "[harness] li: %zu rays in %.3f s\n", rays.size(), now_s() - t0);
    fclose(o);
}

// Li of whichever volume integrator the scene file names ("single", "emission" or "photonvolume"), one RNG(seed + i) per ray,
// the scatter sample taken from the ray record: the golden of SURVEY.md 8(f)-4's SingleScatteringIntegrator / EmissionIntegrator
// where the volume integrator of the scene keeps its two 1-D samples (valid once a Sample was built for it)
static void volint_offsets(VolumeIntegrator *vi, int *scat, int *tau) {
    if (SingleScatteringIntegrator *s = dynamic_cast<SingleScatteringIntegrator *>(vi)) { *scat = s->scatterSampleOffset; *tau = s->tauSampleOffset; }
    else if (EmissionIntegrator *e = dynamic_cast<EmissionIntegrator *>(vi)) { *scat = e->scatterSampleOffset; *tau = e->tauSampleOffset; }
    else if (PhotonVolumeIntegrator *v = dynamic_cast<PhotonVolumeIntegrator *>(vi)) { *scat = v->scatterSampleOffset; *tau = v->tauSampleOffset; }
    else { fprintf(stderr, "unknown volume integrator\n"); exit(4); }
}
static void vli(SamplerRenderer *ren, const Scene *scene, const std::string &rfn, uint32_t seed, const std::string &ofn) {
    std::vector<pv_ray> rays = read_rays(rfn);
    int scat = -1, tau = -1;
    Sample *sample = new Sample(ren->sampler, ren->surfaceIntegrator, ren->volumeIntegrator, scene);
    volint_offsets(ren->volumeIntegrator, &scat, &tau);
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVLI0001", rays.size());
    MemoryArena arena;
    for (size_t i = 0; i < rays.size(); ++i) {
        RayDifferential r(to_ray(rays[i]));
        RNG rng(seed + (uint32_t)i);
        sample->oneD[scat][0] = rays[i].u_scatter;
        sample->oneD[tau][0] = 0.5f;
        Spectrum T(1.f);
        Spectrum L = ren->volumeIntegrator->Li(scene, ren, r, sample, rng, &T, arena);
        wr(o, L.c, 30); wr(o, T.c, 30);
        arena.FreeAll();
    }
    fclose(o);
}

static void transmittance(SamplerRenderer *ren, const Scene *scene, const std::string &rfn, uint32_t seed, const std::string &ofn) {
    // sample == NULL branch of photonvolume.cpp:24-27: step = 4*stepSize, offset = rng.RandomFloat().
    // The offset used is written next to T so the caller can replay it.
    std::vector<pv_ray> rays = read_rays(rfn);
    FILE *o = xopen(ofn, "wb");
    write_header(o, "PVTR0001", rays.size());
    MemoryArena arena;
    for (size_t i = 0; i < rays.size(); ++i) {
        RayDifferential r(to_ray(rays[i]));
        RNG rng(seed + (uint32_t)i), rng2(seed + (uint32_t)i);
        float u = rng2.RandomFloat();
        Spectrum T = ren->volumeIntegrator->Transmittance(scene, ren, r, NULL, rng, arena);
        wr(o, &u, 1); wr(o, T.c, 30);
    }
    fclose(o);
}


// ---- the reference's own multithreaded path (core/parallel.cpp task system) over a ray file: what bench.py --impl reference times
class LiTask : public Task {
public:
    LiTask(SamplerRenderer *r, const Scene *sc, const std::vector<pv_ray> *ry, size_t b, size_t e, uint32_t sd, float *o, Sample *orig)
        : ren(r), scene(sc), rays(ry), begin(b), end(e), seed(sd), out(o), origSample(orig) {}
    void Run() {
        VolumeIntegrator *vi = ren->volumeIntegrator;          // photonvolume, single or emission: whichever the scene file names
        int scat = -1, tau = -1;
        volint_offsets(vi, &scat, &tau);
        Sample *sample = origSample->Duplicate(1);
        MemoryArena arena;
        for (size_t i = begin; i < end; ++i) {
            RayDifferential r(to_ray((*rays)[i]));
            RNG rng(seed + (uint32_t)i);
            sample->oneD[scat][0] = (*rays)[i].u_scatter;
            sample->oneD[tau][0] = 0.5f;
            Spectrum T(1.f);
            Spectrum L = vi->Li(scene, ren, r, sample, rng, &T, arena);
            memcpy(out + 60 * i, L.c, 30 * sizeof(float)); memcpy(out + 60 * i + 30, T.c, 30 * sizeof(float));
            arena.FreeAll();
        }
        delete[] sample;
    }
    SamplerRenderer *ren; const Scene *scene; const std::vector<pv_ray> *rays; size_t begin, end; uint32_t seed; float *out; Sample *origSample;
};
static void li_parallel(SamplerRenderer *ren, const Scene *scene, const std::string &rfn, uint32_t seed, const std::string &ofn) {
    std::vector<pv_ray> rays = read_rays(rfn);
    Sample *sample = new Sample(ren->sampler, ren->surfaceIntegrator, ren->volumeIntegrator, scene);
    std::vector<float> out(60 * rays.size());
    std::vector<Task *> tasks;
    const size_t chunk = 64;
    for (size_t b = 0; b < rays.size(); b += chunk)
        tasks.push_back(new LiTask(ren, scene, &rays, b, std::min(rays.size(), b + chunk), seed, out.data(), sample));
    double t0 = now_s();
    EnqueueTasks(tasks);
    WaitForAllTasks();
    double dt = now_s() - t0;
    for (size_t i = 0; i < tasks.size(); ++i) delete tasks[i];
    fprintf(stderr, "[harness] li-parallel: %zu rays in %.6f s on %d cores\n", rays.size(), dt, NumSystemCores());
    if (ofn != "-") {
        FILE *o = xopen(ofn, "wb");
        write_header(o, "PVLI0001", rays.size());
        wr(o, out.data(), out.size());
        fclose(o);
    }
}
// replace the parsed (small) density grid by an n^3 grid read from a raw float file: the 256^3 grid of config 3 would be
// a 120 MB text block in the .pbrt file
static void swap_grid(Scene *scene, int n, const std::string &fn) {
    VolumeGridDensity *gd = dynamic_cast<VolumeGridDensity *>(scene->volumeRegion);
    if (!gd) { fprintf(stderr, "--grid-file needs a volumegrid scene\n"); exit(4); }
    std::vector<float> d((size_t)n * n * n);
    FILE *f = xopen(fn, "rb"); rd(f, d.data(), d.size()); fclose(f);
    Transform v2w = Inverse(gd->WorldToVolume);
    VolumeGridDensity *ng = new VolumeGridDensity(gd->sig_a, gd->sig_s, gd->g, gd->le, gd->extent, v2w, n, n, n, d.data());
    scene->volumeRegion = ng;
    scene->bound = Union(scene->aggregate->WorldBound(), ng->WorldBound());
    fprintf(stderr, "[harness] density grid replaced by %d^3 from %s\n", n, fn.c_str());
}

// ---------------------------------------------------------------- WorldEnd hook
void pbrtWorldEnd() {
    VERIFY_WORLD("WorldEnd");
    Renderer *renderer = renderOptions->MakeRenderer();
    Scene *scene = renderOptions->MakeScene();
    SamplerRenderer *sr = dynamic_cast<SamplerRenderer *>(renderer);
    if (!scene || !sr) { fprintf(stderr, "harness needs the sampler renderer\n"); g_rc = 4; return; }
    PhotonShooter *sh = sr->photonShooter;
    PhotonVolumeIntegrator *vi = dynamic_cast<PhotonVolumeIntegrator *>(sr->volumeIntegrator);
    for (size_t i = 0; i < g_ops.size(); ++i) {
        const std::string &op = g_ops[i];
        #define ARG(k) (i + (k) < g_ops.size() ? g_ops[i + (k)] : (fprintf(stderr, "missing arg for %s\n", op.c_str()), exit(2), g_ops[0]))
        if (op == "--export-area-lights") {
            // scene with DiffuseAreaLights over triangle meshes: <scn> keeps a placeholder in each such light's slot, <side> lists per
            // area light its slot, Lemit and the triangles of its ShapeSet in the ShapeSet's own (refine) order -- the order the
            // area CDF samples by (core/light.cpp:114-137).  PVAREA01: n; per light: slot u32, n_tris u32, flags u32 (1 = reverse
            // orientation, 2 = transform swaps handedness), Lemit f32[30], vertices f32[9 * n_tris] (world space).
            PvHostScene hs; std::string err;
            if (!pv_export_scene(scene, hs, err, false, true)) { fprintf(stderr, "%s\n", err.c_str()); exit(4); }
            if (!pv_write_scene_file(hs, ARG(1))) exit(3);
            FILE *f = xopen(ARG(2), "wb");
            uint64_t n = 0;
            for (size_t l = 0; l < scene->lights.size(); ++l) if (dynamic_cast<DiffuseAreaLight *>(scene->lights[l])) ++n;
            write_header(f, "PVAREA01", n);
            for (size_t l = 0; l < scene->lights.size(); ++l) {
                DiffuseAreaLight *al = dynamic_cast<DiffuseAreaLight *>(scene->lights[l]);
                if (!al) continue;
                uint32_t slot = (uint32_t)l, nt = (uint32_t)al->shapeSet->shapes.size(), flags = 0;
                std::vector<float> verts;
                for (uint32_t k = 0; k < nt; ++k) {
                    const Triangle *t = dynamic_cast<const Triangle *>(al->shapeSet->shapes[k].GetPtr());
                    if (!t) { fprintf(stderr, "--export-area-lights: only triangle-mesh area lights\n"); exit(4); }
                    flags = (t->ReverseOrientation ? 1u : 0u) | (t->TransformSwapsHandedness ? 2u : 0u);
                    for (int c = 0; c < 3; ++c) { const Point &q = t->mesh->p[t->v[c]]; verts.push_back(q.x); verts.push_back(q.y); verts.push_back(q.z); }
                }
                wr(f, &slot, 1); wr(f, &nt, 1); wr(f, &flags, 1); wr(f, al->Lemit.c, 30); wr(f, verts.data(), verts.size());
            }
            fclose(f);
            i += 2;
        }
        else if (op == "--export-regions") {
            // an AggregateVolume (several Volume statements): the whole scene once per region, with that region as its only medium
            // -> <prefix>.<i>.scn; the oracle takes region 0's file as the scene and the media of the others on the side
            AggregateVolume *agg = dynamic_cast<AggregateVolume *>(scene->volumeRegion);
            if (!agg) { fprintf(stderr, "--export-regions needs a scene with more than one Volume\n"); exit(4); }
            Scene *sc = const_cast<Scene *>(scene);
            for (size_t k = 0; k < agg->regions.size(); ++k) {
                sc->volumeRegion = agg->regions[k];
                PvHostScene hs; std::string err;
                if (!pv_export_scene(sc, hs, err)) { fprintf(stderr, "%s\n", err.c_str()); exit(4); }
                std::ostringstream fn; fn << ARG(1) << "." << k << ".scn";
                if (!pv_write_scene_file(hs, fn.str())) exit(3);
            }
            sc->volumeRegion = agg;
            i += 1;
        }
        else if (op == "--export-medium") { PvHostScene hs; std::string err; if (!pv_export_scene(scene, hs, err, true)) { fprintf(stderr, "%s\n", err.c_str()); exit(4); } if (!pv_write_scene_file(hs, ARG(1))) exit(3); i += 1; }
        else if (op == "--export-scene") { PvHostScene hs; std::string err; if (!pv_export_scene(scene, hs, err)) { fprintf(stderr, "%s\n", err.c_str()); exit(4); } if (!pv_write_scene_file(hs, ARG(1))) exit(3); i += 1; }
        else if (op == "--shoot") { shoot(sh, scene, sr->camera, sr); }
        else if (op == "--load-photons") { load_photons(ARG(1)); install_volume_map(sh); i += 1; }
        else if (op == "--dump-photons") { dump_photons(ARG(1)); i += 1; }
        else if (op == "--dump-maps") { dump_maps(ARG(1)); i += 1; }
        else if (op == "--surface-lphoton") { surface_lphoton(sh, ARG(1), ARG(2), ARG(3)); i += 3; }
        else if (op == "--radiance-nearest") { radiance_nearest(ARG(1), ARG(2)); i += 2; }
        else if (op == "--knn") { knn(sh, ARG(1), atoi(ARG(2).c_str()), (float)atof(ARG(3).c_str()), ARG(4)); i += 4; }
        else if (op == "--lphoton") { lphoton(vi, scene, ARG(1), ARG(2)); i += 2; }
        else if (op == "--intersect") { intersect(scene, ARG(1), ARG(2)); i += 2; }
        else if (op == "--li") { li(sr, scene, ARG(1), (uint32_t)strtoul(ARG(2).c_str(), NULL, 0), ARG(3)); i += 3; }
        else if (op == "--vli") { vli(sr, scene, ARG(1), (uint32_t)strtoul(ARG(2).c_str(), NULL, 0), ARG(3)); i += 3; }
        else if (op == "--li-parallel") { li_parallel(sr, scene, ARG(1), (uint32_t)strtoul(ARG(2).c_str(), NULL, 0), ARG(3)); i += 3; }
        else if (op == "--grid-file") { swap_grid(scene, atoi(ARG(1).c_str()), ARG(2)); i += 2; }
        else if (op == "--transmittance") { transmittance(sr, scene, ARG(1), (uint32_t)strtoul(ARG(2).c_str(), NULL, 0), ARG(3)); i += 3; }
        else if (op == "--stats") {
            FILE *f = xopen(ARG(1), "w");
            fprintf(f, "{\"nshot\": %u, \"volume_photons\": %zu, \"shoot_seconds\": %.6f, \"cores\": %d, \"caustic_paths\": %d, "
                    "\"indirect_paths\": %d, \"direct_paths\": %d, \"volume_paths\": %d, \"nlookup\": %u, \"maxdist2\": %.9g, \"final_gather\": %d}\n",
                    g_nshot, g_photons.size(), g_shoot_seconds, NumSystemCores(), g_paths[0], g_paths[1], g_paths[2], g_paths[3],
                    sh->nLookup, sh->maxDistSquared, (int)sh->finalGather);
            fclose(f); i += 1;
        }
        else if (op == "--render") { renderer->Render(scene); }
        else { fprintf(stderr, "unknown op %s\n", op.c_str()); g_rc = 2; }
        #undef ARG
    }
    TasksCleanup();
    // like the reference's pbrtWorldEnd (core/api.cpp:1176-1193) minus the render
    graphicsState = GraphicsState();
    transformCache.Clear();
    currentApiState = STATE_OPTIONS_BLOCK;
    for (int i = 0; i < MAX_TRANSFORMS; ++i) curTransform[i] = Transform();
    activeTransformBits = ALL_TRANSFORMS_BITS;
    namedCoordinateSystems.erase(namedCoordinateSystems.begin(), namedCoordinateSystems.end());
}

int main(int argc, char *argv[]) {
    if (argc < 2) {
        fprintf(stderr, "usage: ref_harness scene.pbrt [--ncores N] ops...\n");
        return 2;
    }
    Options options;
    options.nCores = 1;
    options.quiet = true;
    std::string scene = argv[1];
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "--ncores") && i + 1 < argc) options.nCores = atoi(argv[++i]);
        else g_ops.push_back(argv[i]);
    }
    pbrtInit(options);
    if (!ParseFile(scene)) { fprintf(stderr, "could not parse %s\n", scene.c_str()); return 2; }
    pbrtCleanup();
    return g_rc;
}
