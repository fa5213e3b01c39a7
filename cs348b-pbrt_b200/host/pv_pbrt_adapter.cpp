// pv_pbrt_adapter.cpp -- the host side of the drop-in: the reference's own classes, re-implemented over the C ABI.
//
// Linked INTO the reference (piwell/CS348B-pbrt) in place of five of its translation units; every class keeps the
// declaration of the reference's header, so core/api.cpp (the name-string "plugin registry", :572-586, :1221-1288),
// the parser, cameras, samplers, surface integrators and the film are used unchanged:
//
//   replaced TU                          what this file provides instead
//   core/photonshooter.cpp   (:457-526)  PhotonShooter::Preprocess -> pv_set_scene + pv_shoot_maps (all photon maps in one
//                                        GPU pass) + pv_radiance_photons + pv_build.  The caustic / indirect / radiance maps the
//                                        untouched PhotonIntegrator reads (integrators/photonmap.cpp:159-164) are rebuilt as
//                                        the reference's KdTree<> objects from the device's photon lists.  The rest of that TU
//                                        is still the reference's, compiled with -DPreprocess=RefPreprocess: its CPU pass stays
//                                        reachable for scenes off this path (and, in a -DPV_WITH_CPU_ROUTES debugging build only,
//                                        under PV_SURFACE_MAPS=cpu).
//   integrators/photonvolume.cpp         PhotonVolumeIntegrator::{RequestSamples, Transmittance, Li} and
//                                        CreatePhotonVolumeIntegrator (same .pbrt parameters, :224-229).
//   integrators/single.cpp, emission.cpp SingleScatteringIntegrator / EmissionIntegrator: Li over pv_volume_li.
//   renderers/samplerrenderer.cpp        SamplerRenderer with a BATCHED Render: image tiles run the unchanged camera /
//                                        sampler / surface-integrator code on the reference's pthread pool
//                                        (core/parallel.cpp) and queue their camera rays; the volume term of the whole
//                                        frame is ONE pv_gather call; samples are then added to the unchanged Film.
//
// Nothing here computes radiance on the CPU: Li without a CUDA device is an error (Severe), not a fallback.
// PhotonVolumeIntegrator::Transmittance, which the surface integrators call one ray at a time, stays the reference's
// three lines over VolumeRegion::tau (core/volume.cpp) -- it is not worth a device round trip per shadow ray; the GPU
// uses its own batch form (pv_transmittance and the in-kernel marches).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <vector>
#include <string>
#include <map>
#include <algorithm>
#include <sys/time.h>
#include <mutex>
#include <condition_variable>
#include <memory>
#include <chrono>
#include <thread>

#define private public
#define protected public
#include "stdafx.h"
#include "core/pbrt.h"
#include "core/scene.h"
#include "core/light.h"
#include "core/sampler.h"
#include "core/camera.h"
#include "core/film.h"
#include "film/image.h"
#include "core/filter.h"
#include "core/intersection.h"
#include "core/paramset.h"
#include "core/montecarlo.h"
#include "core/progressreporter.h"
#include "core/photonshooter.h"
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
#include "integrators/single.h"
#include "integrators/emission.h"
#include "integrators/photonmap.h"
#include "renderers/samplerrenderer.h"
#undef private
#undef protected

#include "../../include/pv.h"
#include "pv_export.inl"

// PhotonShooter::Preprocess of the reference's own TU, renamed at compile time (see host/Makefile); called through its
// Itanium-ABI name because the class declaration cannot be extended.
extern "C" void _ZN13PhotonShooter13RefPreprocessEPK5ScenePK6CameraPK8Renderer(PhotonShooter *, const Scene *, const Camera *,
                                                                              const Renderer *);

static void pv_fill_ray(const RayDifferential &ray, float u_scatter, pv_ray *r);

namespace {
struct PvBridge {
    pv_ctx *ctx;
    PvHostScene scene;
    float stepsize, maxdist;
    uint32_t nused;
    uint64_t seed;
    uint64_t n_indirect;
    bool ready;
    int volint;                   // which Li the device runs: -1 photonvolume (pv_gather), else PV_VOLINT_* (pv_volume_li)
    // final gathering runs on a context of its own (scene + radiance photons + their grid): rays spawned by specular bounces
    // call the volume integrator (pv_gather on `ctx`, grid on the volume map) while primary hits are still being shaded
    pv_ctx *fg_ctx;
    // PV_DEVICES="0,1,2,3": contexts on the other GPUs of the box, each holding the scene and a replica of the volume photon map
    // (replicate the map, shard the camera rays).  Empty unless PV_DEVICES names more than one device.
    std::vector<pv_ctx *> replicas;
    std::vector<float> rad_pos, rad_nrm, rad_Lo;
    // the caustic and indirect maps as the device pass left them (planes), for the LPhoton lookups of primary hits; and the
    // contexts that hold them with their grids (one lookup grid per context)
    std::vector<float> surf_pos[2], surf_wi[2], surf_alpha[2];
    pv_ctx *surf_ctx[2];
    PvBridge() : ctx(NULL), stepsize(1.f), maxdist(.1f), nused(250), seed(0), n_indirect(0), ready(false), volint(-1), fg_ctx(NULL) { surf_ctx[0] = surf_ctx[1] = NULL; }
};
PvBridge g_pv;

double now_s() { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + 1e-6 * tv.tv_usec; }

void pv_fail(const char *what, int rc) {
    Severe("%s failed (%d): %s", what, rc, pv_last_error(g_pv.ctx));
}

// The scene BVH built on the device (pv_build_bvh, csrc/pv_lbvh.cu) for scenes whose aggregate holds no LinearBVHNode array
// (Accelerator "kdtree", pbrt's own default, and "grid"), or for every scene under PV_BVH=gpu.  The reference's per-ray CPU code
// (surface integrator) keeps using the scene's own aggregate; the device kernels use this tree.
struct PvDeviceBvh : PvBvhBuilder {
    bool build(const float *bounds, uint32_t n, uint32_t max_prims, std::vector<pv_bvh_node> &nodes, std::vector<uint32_t> &order,
               std::string &err) {
        nodes.resize(n ? 2 * (size_t)n - 1 : 0); order.resize(n);
        uint32_t n_nodes = 0; float ms = 0.f;
        double t0 = now_s();
        int rc = pv_build_bvh(g_pv.ctx, bounds, n, max_prims, nodes.data(), (uint32_t)nodes.size(), &n_nodes, order.data(), &ms);
        if (rc) { err = std::string("pv_build_bvh: ") + pv_last_error(g_pv.ctx); return false; }
        nodes.resize(n_nodes);
        fprintf(stderr, "[pv] scene BVH built on the GPU: %u primitives -> %u nodes (leaves of <= %u), %.2f ms of kernels, %.1f ms with copies\n",
                n, n_nodes, max_prims, ms, 1e3 * (now_s() - t0));
        return true;
    }
};
std::vector<int> pv_device_list();
// export + pv_create + pv_set_scene, the common start of every device path
bool pv_export_and_set(const Scene *scene, std::string &err, bool medium_only = false) {
    if (!g_pv.ctx) {
        int rc = pv_create(&g_pv.ctx, pv_device_list()[0]);
        if (rc) Severe("pv_create failed (%d): %s", rc, pv_last_error(NULL));
    }
    PvDeviceBvh builder;
    const char *mode = getenv("PV_BVH");
    builder.force = mode && !strcmp(mode, "gpu");
    if (!pv_export_scene(scene, g_pv.scene, err, medium_only, false, &builder)) return false;
    int rc = pv_set_scene(g_pv.ctx, &g_pv.scene.desc);
    if (rc) pv_fail("pv_set_scene", rc);
    return true;
}

// The three VolumeIntegrator plugins whose Li runs on the device, seen through one pair of glasses.
struct PvVolInt { int kind; float stepSize; int nUsed; float maxDist; int scatterSampleOffset; };
bool pv_volint_of(const VolumeIntegrator *v, PvVolInt *o) {
    if (const PhotonVolumeIntegrator *p = dynamic_cast<const PhotonVolumeIntegrator *>(v)) {
        o->kind = -1; o->stepSize = p->stepSize; o->nUsed = p->nUsed; o->maxDist = p->maxDist; o->scatterSampleOffset = p->scatterSampleOffset;
    } else if (const SingleScatteringIntegrator *s1 = dynamic_cast<const SingleScatteringIntegrator *>(v)) {
        o->kind = PV_VOLINT_SINGLE; o->stepSize = s1->stepSize; o->nUsed = 0; o->maxDist = 0.f; o->scatterSampleOffset = s1->scatterSampleOffset;
    } else if (const EmissionIntegrator *e = dynamic_cast<const EmissionIntegrator *>(v)) {
        o->kind = PV_VOLINT_EMISSION; o->stepSize = e->stepSize; o->nUsed = 0; o->maxDist = 0.f; o->scatterSampleOffset = e->scatterSampleOffset;
    } else return false;
    return true;
}
// Li of a batch of rays with whichever of the three the scene file chose
int pv_volume_term_on(pv_ctx *ctx, const pv_ray *rays, size_t n, const pv_gather_params *prm, float *L, float *T) {
    if (g_pv.volint < 0) return pv_gather(ctx, rays, n, prm, L, T);
    return pv_volume_li(ctx, g_pv.volint, rays, n, prm, L, T);
}
int pv_volume_term(const pv_ray *rays, size_t n, const pv_gather_params *prm, float *L, float *T) {
    return pv_volume_term_on(g_pv.ctx, rays, n, prm, L, T);
}

// ---- several GPUs inside the drop-in (off unless PV_DEVICES lists more than one device) -----
// The scene and the volume photon map are replicated (the whole map is a few GB at most); the camera rays of
// the frame are cut into one contiguous run per device, each with its global ray index as Philox stream base, so the image does
// not depend on the number of devices.  Shooting, final gathering and the secondary rays of specular bounces stay on the first.
std::vector<int> pv_device_list() {
    std::vector<int> d;
    if (const char *e = getenv("PV_DEVICES"))
        for (const char *p = e; *p;) { char *q; long v = strtol(p, &q, 10); if (q == p) break; d.push_back((int)v); p = *q == ',' ? q + 1 : q; }
    if (d.empty()) { const char *e = getenv("PV_DEVICE"); d.push_back(e ? atoi(e) : 0); }
    return d;
}
void pv_replicate(bool with_map, float maxdist, uint32_t nused) {
    const std::vector<int> dev = pv_device_list();
    for (size_t i = 0; i < g_pv.replicas.size(); ++i) pv_destroy(g_pv.replicas[i]);
    g_pv.replicas.clear();
    if (dev.size() < 2) return;
    for (size_t i = 1; i < dev.size(); ++i) {
        pv_ctx *c = NULL;
        int rc = pv_create(&c, dev[i]);
        if (rc) Severe("pv_create on device %d failed (%d): %s", dev[i], rc, pv_last_error(NULL));
        rc = pv_set_scene(c, &g_pv.scene.desc);
        if (rc) Severe("scene on device %d failed (%d): %s", dev[i], rc, pv_last_error(c));
        g_pv.replicas.push_back(c);
    }
    float ms = 0.f;
    if (with_map) {
        // the map goes GPU to GPU over NVLink: one NCCL communicator over the contexts, one grouped broadcast of the four SoA
        // planes (pv_broadcast_photons, csrc/pv_comm.cu); every replica then builds its own grid
        std::vector<pv_ctx *> all(1, g_pv.ctx);
        all.insert(all.end(), g_pv.replicas.begin(), g_pv.replicas.end());
        pv_comm_destroy(g_pv.ctx);                        // a second Preprocess of the same process starts from a fresh communicator
        int rc = pv_comm_init_all(all.data(), (int)all.size());
        if (rc) Severe("pv_comm_init_all over %zu devices failed (%d): %s", all.size(), rc, pv_last_error(g_pv.ctx));
        rc = pv_broadcast_photons(all.data(), (int)all.size(), 0, &ms);
        if (rc) Severe("pv_broadcast_photons failed (%d): %s", rc, pv_last_error(g_pv.ctx));
        for (size_t i = 0; i < g_pv.replicas.size(); ++i) {
            rc = pv_build(g_pv.replicas[i], maxdist, nused);
            if (rc) Severe("map build on device %d failed (%d): %s", dev[i + 1], rc, pv_last_error(g_pv.replicas[i]));
        }
    }
    fprintf(stderr, "[pv] scene%s replicated on %zu more device(s)", with_map ? " and volume photon map" : "", g_pv.replicas.size());
    if (with_map) fprintf(stderr, " (NCCL broadcast %.2f ms)", ms);
    fprintf(stderr, "\n");
}
// the volume term of a whole frame: one call per device, concurrently
int pv_volume_term_frame(const pv_ray *rays, size_t n, const pv_gather_params *prm, float *L, float *T) {
    const size_t G = 1 + g_pv.replicas.size();
    if (G == 1 || n < 4096 * G) return pv_volume_term(rays, n, prm, L, T);
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    for (size_t d = 0; d < G; ++d) {
        const size_t a = n * d / G, b = n * (d + 1) / G;
        pv_ctx *c = d == 0 ? g_pv.ctx : g_pv.replicas[d - 1];
        th.push_back(std::thread([=, &rcs]() {
            pv_gather_params p = *prm; p.ray_index_base = prm->ray_index_base + a;
            rcs[d] = pv_volume_term_on(c, rays + a, b - a, &p, L + a * PV_NSPEC, T + a * PV_NSPEC);
        }));
    }
    for (size_t d = 0; d < G; ++d) th[d].join();
    for (size_t d = 0; d < G; ++d) if (rcs[d]) { if (d) Error("device call on replica %zu failed: %s", d, pv_last_error(g_pv.replicas[d - 1])); return rcs[d]; }
    return 0;
}

struct PvRecord {                 // one camera sample waiting for its volume term
    float imageX, imageY, rayWeight;
    Spectrum Ls;                  // surface radiance (before the volume transmittance)
    pv_ray ray;
};
std::vector<std::vector<PvRecord> > *g_records = NULL;

// ---- final gathering of PRIMARY hits as a wavefront (integrators/photonmap.cpp:196-312) ----------------------------
// The reference's PhotonIntegrator::Li traces 2 x gatherSamples rays per shaded point and, for each, looks the nearest radiance
// photon up and marches the medium for the transmittance -- one ray at a time on the CPU; with final gathering on this is
// nearly all of a render's time.  Here the task that shades a camera sample only GENERATES those rays (the reference's own
// BSDF / sampling code and its MIS weights) and records, per ray, the spectrum C its radiance is multiplied with; the rays of
// many samples go to the GPU as one pv_final_gather batch; then Ls += sum C * Lindir.  Everything else of Li (emission, direct
// lighting, the caustic estimate, specular bounces) is the reference's own code, run through a clone of the integrator whose
// shooter has no indirect map (so that Li adds neither the final gather nor the indirect estimate); rays spawned by specular
// bounces re-enter through SamplerRenderer::Li and use the unmodified integrator, final gathering included.
struct PvGatherRay { uint32_t rec; int tslot; Spectrum C; pv_ray ray; };   // tslot: see PvSecondary
// One LPhoton lookup of a primary hit (photonmap.cpp:62-108, diffuse branch): where, the faceforwarded shading normal, and the
// two reflectances the sums are multiplied with afterwards.
struct PvLookup { uint32_t rec; int tslot; float p[3], nf[3]; Spectrum rr, rt; };
// One light sample of the direct lighting of a primary hit (EstimateDirect, core/integrator.cpp:137-163, delta light): the
// VisibilityTester's segment, the jitter of the transmittance march (the rng.RandomFloat() of photonvolume.cpp:26) and the
// spectrum C = f * Li * |wi . n| / pdf the visibility V * Tr is multiplied with.
struct PvShadow { uint32_t rec; int tslot; Spectrum C; pv_ray ray; float u; };
// A ray behind a specular bounce (SpecularReflect / SpecularTransmit): its volume term Lv, T is ONE ray of the group's single
// pv_gather_indexed call instead of a blocking call of its own.  W = the product of the bounce factors f * |wi . n| / pdf up to
// this ray (no transmittances), parent = the secondary ray of the same task this one was spawned behind (-1: behind a primary hit).
// What the ray contributes: W * P(parent) * Lv with P(j) = P(parent(j)) * T(j) the product of the volume transmittances along the
// chain, P(-1) = 1.  The terms queued at the hit this ray ends in carry tslot = the ray's number: their weights are multiplied
// with P(tslot) once the call is back.
struct PvSecondary { uint32_t rec; int parent; Spectrum W; pv_ray ray; uint64_t index; };
struct PvFinalGather {
    PhotonIntegrator *full, *primary;      // the scene's integrator, and its clone without the maps whose terms run on the device
    bool gather_rays;                      // final gathering of primary hits as one pv_final_gather per group
    bool lookup[2];                        // LPhoton of primary hits on the caustic map (:179) / on the indirect map when final gathering is off (:308)
    bool direct;                           // direct lighting of primary hits: shadow rays + their transmittance as one device batch per group
    bool defer;                            // rays behind specular bounces are deferred to one call per group (PvSecondary); PV_DEFER_SECONDARY=0: blocking batched calls (A/B)
    std::vector<std::vector<PvShadow> > shadows;      // per render task
    uint64_t total_shadows; double shadow_seconds;
    std::vector<std::vector<PvSecondary> > secondary; // per render task (direct mode: the rays behind specular bounces are deferred too)
    uint64_t total_secondary, secondary_calls; double secondary_seconds;
    int scatter_offset; pv_gather_params vol_prm;     // of the scene's PhotonVolumeIntegrator
    std::vector<std::vector<PvGatherRay> > rays;      // per render task
    std::vector<std::vector<PvLookup> > lookups;      // per render task; caustic and indirect lookups of a hit share the record
    uint64_t next_index, total_rays, total_lookups; double gpu_seconds, lookup_seconds;
    PvFinalGather() : full(NULL), primary(NULL), gather_rays(false), direct(false), defer(true), total_shadows(0), shadow_seconds(0), total_secondary(0), secondary_calls(0), secondary_seconds(0), scatter_offset(0), next_index(0), total_rays(0), total_lookups(0), gpu_seconds(0), lookup_seconds(0) {
        lookup[0] = lookup[1] = false;
    }
};
PvFinalGather *g_fg = NULL;

// The sampling half of the final gather for one shaded point: which rays, with which weights (photonmap.cpp:204-312).
void pv_queue_final_gather(const PhotonIntegrator *pi, const RayDifferential &ray, const Intersection &isect, const Sample *sample,
                           MemoryArena &arena, uint32_t rec, std::vector<PvGatherRay> &out) {
    KdTree<Photon> *indirectMap = pi->photonShooter->indirectMap;
    BSDF *bsdf = isect.GetBSDF(ray, arena);
    const BxDFType nonSpecular = BxDFType(BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_DIFFUSE | BSDF_GLOSSY);
    if (bsdf->NumComponents(nonSpecular) == 0) return;
    const Point &p = bsdf->dgShading.p;
    const Normal &n = bsdf->dgShading.nn;
    const Vector wo = -ray.d;
    // directions of the 50 nearest indirect photons: the importance function of the second half of the rays (:204-219)
    const uint32_t nDirs = 50;
    ClosePhoton close[nDirs];
    PhotonProcess proc(nDirs, close);
    float searchDist2 = pi->maxDistSquared;
    while (proc.nFound < nDirs) {
        float md2 = searchDist2;
        proc.nFound = 0;
        indirectMap->Lookup(p, proc, md2);
        searchDist2 *= 2.f;
    }
    Vector dirs[nDirs];
    for (uint32_t i = 0; i < nDirs; ++i) dirs[i] = close[i].photon->wi;
    const int gs = pi->gatherSamples;
    const float cosGA = pi->cosGatherAngle, conePdf = UniformConePdf(cosGA);
    PvGatherRay g; memset(&g.ray, 0, sizeof(g.ray)); g.rec = rec; g.tslot = -1;
    g.ray.o[0] = p.x; g.ray.o[1] = p.y; g.ray.o[2] = p.z; g.ray.mint = isect.rayEpsilon; g.ray.maxt = INFINITY; g.ray.time = ray.time;
    for (int half = 0; half < 2; ++half)
        for (int i = 0; i < gs; ++i) {
            Vector wi; float pdf = 0.f; Spectrum fr;
            if (half == 0) {                                                   // BSDF-sampled ray (:223-229)
                BSDFSample bs(sample, pi->bsdfGatherSampleOffsets, i);
                fr = bsdf->Sample_f(wo, &wi, bs, &pdf, BxDFType(BSDF_ALL & ~BSDF_SPECULAR));
                if (fr.IsBlack() || pdf == 0.f) continue;
            } else {                                                           // ray in a cone around a photon direction (:264-275)
                BSDFSample gsamp(sample, pi->indirGatherSampleOffsets, i);
                int photonNum = min((int)nDirs - 1, Floor2Int(gsamp.uComponent * nDirs));
                Vector vx, vy;
                CoordinateSystem(dirs[photonNum], &vx, &vy);
                wi = UniformSampleCone(gsamp.uDir[0], gsamp.uDir[1], cosGA, vx, vy, dirs[photonNum]);
                fr = bsdf->f(wo, wi);
                if (fr.IsBlack()) continue;
            }
            float photonPdf = 0.f;                                             // pdf of the photon-direction strategy for wi (:246-252)
            for (uint32_t j = 0; j < nDirs; ++j) if (Dot(dirs[j], wi) > .999f * cosGA) photonPdf += conePdf;
            photonPdf /= nDirs;
            float scale;
            if (half == 0) scale = AbsDot(wi, n) * PowerHeuristic(gs, pdf, gs, photonPdf) / pdf;                       // :253-254
            else scale = AbsDot(wi, n) * PowerHeuristic(gs, photonPdf, gs, bsdf->Pdf(wo, wi)) / photonPdf;            // :299-301
            g.C = fr * (scale / gs);
            g.ray.d[0] = wi.x; g.ray.d[1] = wi.y; g.ray.d[2] = wi.z;
            out.push_back(g);
        }
}
// The host half of LPhoton for one shaded point (photonmap.cpp:62-108).  On this path a surface is Lambertian or purely specular
// (any other material stops the drop-in before a photon is shot), so a point with non-specular components takes the DIFFUSE
// branch: L = Lr * rho_r / pi + Lt * rho_t / pi, with Lr / Lt the kernel-weighted flux of the nLookup nearest photons arriving on
// either side of Nf -- those two sums are pv_surface_lphoton; the reflectances are the reference's own BSDF::rho (which draws its
// 2 x 36 stratified samples from the task's RNG exactly as LPhoton does, whatever the BxDFs then do with them).
bool pv_queue_lookup(const RayDifferential &ray, const Intersection &isect, RNG &rng, MemoryArena &arena, uint32_t rec,
                     std::vector<PvLookup> &out) {
    BSDF *bsdf = isect.GetBSDF(ray, arena);
    const BxDFType nonSpecular = BxDFType(BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_DIFFUSE | BSDF_GLOSSY);
    if (bsdf->NumComponents(nonSpecular) == 0) return true;                    // LPhoton adds nothing (:68)
    if (bsdf->NumComponents(BxDFType(BSDF_REFLECTION | BSDF_TRANSMISSION | BSDF_GLOSSY)) > 0) return false;   // glossy branch: not on this path
    const Vector wo = -ray.d;
    const Normal Nf = Faceforward(bsdf->dgShading.nn, wo);
    PvLookup q; q.rec = rec; q.tslot = -1;
    q.p[0] = isect.dg.p.x; q.p[1] = isect.dg.p.y; q.p[2] = isect.dg.p.z;
    q.nf[0] = Nf.x; q.nf[1] = Nf.y; q.nf[2] = Nf.z;
    q.rr = bsdf->rho(wo, rng, BSDF_ALL_REFLECTION) * INV_PI;
    q.rt = bsdf->rho(wo, rng, BSDF_ALL_TRANSMISSION) * INV_PI;
    out.push_back(q);
    return true;
}
// PhotonIntegrator::Li for a hit whose every costly term runs on the device (integrators/photonmap.cpp:154-319) -- a primary hit
// (throughput W = 1) or a hit behind specular bounces (W = the product of f * |wi . n| / pdf and of the volume transmittances
// along the chain up to here):
//   emission                       isect.Le(wo), here                                                              (:170)
//   direct lighting                UniformSampleAllLights -> EstimateDirect (core/integrator.cpp:47-79, 137-163) for delta lights:
//                                  Sample_L, the BSDF value and the cosine / pdf factor here with the reference's own objects; the
//                                  shadow ray (VisibilityTester::Unoccluded = !scene->IntersectP) and its transmittance through
//                                  the medium (VisibilityTester::Transmittance -> VolumeIntegrator::Transmittance, sample == NULL:
//                                  step 4 * stepSize, offset rng.RandomFloat()) queued: pv_occluded + pv_transmittance per group
//   caustic / indirect LPhoton     queued (pv_queue_lookup)                                                        (:179, :308)
//   final gathering                queued (pv_queue_final_gather)                                                  (:183-296)
//   specular bounces               the reference's SpecularReflect / SpecularTransmit, here (:310-317).  Their rays re-enter through
//                                  SamplerRenderer::Li, which finds the throughput of the bounce in tl_path and shades the next hit
//                                  with this function again: the recursion only carries Le and the volume term back up, every
//                                  queued term goes straight to the pixel's record with its throughput folded into its weight.
// Used when all lights of the scene are delta lights (an area light's BSDF-sampling half needs the identity of the surface the
// sampled ray hits; such scenes keep the clone route, with the reference's own UniformSampleAllLights).
struct PvPath { bool on; Spectrum W; int task; uint32_t rec; int slot; PvPath() : on(false), W(1.f), task(-1), rec(0), slot(-1) {} };
thread_local PvPath tl_path;

Spectrum pv_hit_li(const SamplerRenderer *r, const Scene *scene, const RayDifferential &ray, const Intersection &isect, const Sample *sample,
                   RNG &rng, MemoryArena &arena, int task, uint32_t rec, const Spectrum &W, int tslot) {
    const PhotonIntegrator *pi = g_fg->full;
    Spectrum L(0.f);
    const Vector wo = -ray.d;
    L += isect.Le(wo);
    BSDF *bsdf = isect.GetBSDF(ray, arena);
    const Point &p = bsdf->dgShading.p;
    const Normal &n = bsdf->dgShading.nn;
    std::vector<PvShadow> &out = g_fg->shadows[task];
    for (uint32_t i = 0; i < scene->lights.size(); ++i) {
        const Light *light = scene->lights[i];
        const int nSamples = pi->lightSampleOffsets ? pi->lightSampleOffsets[i].nSamples : 1;
        for (int j = 0; j < nSamples; ++j) {
            LightSample lightSample = pi->lightSampleOffsets ? LightSample(sample, pi->lightSampleOffsets[i], j) : LightSample(rng);
            Vector wi; float lightPdf; VisibilityTester visibility;
            Spectrum Li = light->Sample_L(p, isect.rayEpsilon, lightSample, ray.time, &wi, &lightPdf, &visibility);
            if (lightPdf > 0.f && !Li.IsBlack()) {
                Spectrum f = bsdf->f(wo, wi, BxDFType(BSDF_ALL & ~BSDF_SPECULAR));
                if (!f.IsBlack()) {
                    PvShadow sh; sh.rec = rec; sh.tslot = tslot;
                    sh.C = W * (f * Li * (AbsDot(wi, n) / lightPdf) / (float)nSamples);
                    pv_fill_ray(RayDifferential(visibility.r), 0.f, &sh.ray);
                    sh.u = rng.RandomFloat();
                    out.push_back(sh);
                }
            }
        }
    }
    if (g_fg->gather_rays) {
        std::vector<PvGatherRay> &gr = g_fg->rays[task];
        const size_t first = gr.size();
        pv_queue_final_gather(pi, ray, isect, sample, arena, rec, gr);
        for (size_t k = first; k < gr.size(); ++k) { gr[k].C *= W; gr[k].tslot = tslot; }
    }
    if (g_fg->lookup[0] || g_fg->lookup[1]) {
        std::vector<PvLookup> &lk = g_fg->lookups[task];
        const size_t first = lk.size();
        if (!pv_queue_lookup(ray, isect, rng, arena, rec, lk))
            Severe("pv: a surface with glossy components reached the device LPhoton (materials other than matte / glass are off this path)");
        for (size_t k = first; k < lk.size(); ++k) { lk[k].rr *= W; lk[k].rt *= W; lk[k].tslot = tslot; }
    }
    if (ray.depth + 1 < pi->maxSpecularDepth) {
        // the factor SpecularReflect / SpecularTransmit will multiply the next hit's radiance with (core/integrator.cpp:184-187,209;
        // 223-226,254), known before they run: a perfectly specular BxDF's Sample_f ignores its sample values
        for (int kind = 0; kind < 2; ++kind) {
            Vector wi; float pdf;
            const BxDFType type = BxDFType((kind == 0 ? BSDF_REFLECTION : BSDF_TRANSMISSION) | BSDF_SPECULAR);
            Spectrum f = bsdf->Sample_f(wo, &wi, BSDFSample(.5f, .5f, .5f), &pdf, type);
            const bool traced = pdf > 0.f && !f.IsBlack() && AbsDot(wi, n) != 0.f;       // else the reference function returns 0 without a ray
            const PvPath saved = tl_path;
            if (traced) { tl_path.on = true; tl_path.W = W * (f * (AbsDot(wi, n) / pdf)); tl_path.task = task; tl_path.rec = rec; tl_path.slot = tslot; }
            L += kind == 0 ? SpecularReflect(ray, bsdf, rng, isect, r, scene, sample, arena)
                           : SpecularTransmit(ray, bsdf, rng, isect, r, scene, sample, arena);
            tl_path = saved;
        }
    }
    return L;
}
}  // namespace

// ------------------------------------------------------------------ PhotonShooter::Preprocess (core/photonshooter.cpp:457-526)
void PhotonShooter::Preprocess(const Scene *scene, const Camera *camera, const Renderer *renderer) {
    if (scene->lights.size() == 0) return;
    const SamplerRenderer *sr = dynamic_cast<const SamplerRenderer *>(renderer);
    PhotonVolumeIntegrator *vi = sr ? dynamic_cast<PhotonVolumeIntegrator *>(sr->volumeIntegrator) : NULL;
    if (!vi || !scene->volumeRegion) {
        // not the photon-volume path: the reference's CPU pass, untouched
        _ZN13PhotonShooter13RefPreprocessEPK5ScenePK6CameraPK8Renderer(this, scene, camera, renderer);
        return;
    }
    std::string err;
    if (!pv_export_and_set(scene, err)) Severe("%s", err.c_str());
    int rc;
    g_pv.stepsize = vi->stepSize; g_pv.maxdist = vi->maxDist; g_pv.nused = (uint32_t)vi->nUsed; g_pv.volint = -1;
    const char *seed = getenv("PV_SEED");
    g_pv.seed = seed ? strtoull(seed, NULL, 0) : 0;
    pv_shoot_params prm; memset(&prm, 0, sizeof(prm));
    prm.stepsize = stepSize; prm.integrator_stepsize = vi->stepSize; prm.max_photon_depth = maxPhotonDepth;
    prm.seed = g_pv.seed; prm.rank = 0; prm.world = 1; prm.time = camera ? camera->shutterOpen : 0.f;
    const bool surface_wanted = nCausticPhotonsWanted + nIndirectPhotonsWanted > 0;
#ifdef PV_WITH_CPU_ROUTES          // debugging build only (host/Makefile CPU_ROUTES=1): PV_SURFACE_MAPS=cpu keeps the reference's CPU pass for the surface maps
    const char *surf_mode = getenv("PV_SURFACE_MAPS");
    const bool surface_gpu = surface_wanted && !(surf_mode && !strcmp(surf_mode, "cpu"));
#else
    const bool surface_gpu = surface_wanted;
#endif
    // The surface maps are read by the unmodified PhotonIntegrator, so the device pass must trace exactly the materials the scene
    // has.  A material the device description cannot hold (anything but Lambertian matte with one Kd and glass) would give wrong
    // caustic / indirect / radiance maps without any sign of it in the image log: refuse instead.
    if (surface_gpu && !g_pv.scene.inexact.empty())
        Severe("pv: the photon maps of this scene cannot be traced on the device: %s (matte with a constant Kd and sigma = 0, and glass, are on this path)",
               g_pv.scene.inexact.c_str());
    if (!surface_gpu && g_pv.scene.has_unknown_specular)
        Severe("pv: the volume photons of this scene cannot be traced on the device: %s", g_pv.scene.inexact.c_str());
    if (surface_gpu) {
        // ---- every map in ONE GPU pass (photonshooter.cpp:147-189, 232-357); the untouched PhotonIntegrator keeps reading
        // causticMap / indirectMap / radianceMap / nCausticPaths / nIndirectPaths (integrators/photonmap.cpp:159-164), so they
        // are rebuilt here as the reference's own KdTree<> objects from what the device returns.
        pv_maps_params mp; memset(&mp, 0, sizeof(mp));
        mp.n_volume_wanted = nVolumePhotonsWanted; mp.n_caustic_wanted = nCausticPhotonsWanted; mp.n_indirect_wanted = nIndirectPhotonsWanted;
        mp.final_gather = finalGather ? 1 : 0;
        pv_maps_stats ms;
        double t0 = now_s();
        rc = pv_shoot_maps(g_pv.ctx, &mp, &prm, &ms);
        if (rc == PV_ENOPHOTONS) Error("Unable to store enough photons.  Giving up.\n");      // photonshooter.cpp:292
        else if (rc) pv_fail("pv_shoot_maps", rc);
        nCausticPaths = (int)ms.n_caustic_paths; nIndirectPaths = (int)ms.n_indirect_paths; nVolumePaths = (int)ms.n_volume_paths;
        g_pv.n_indirect = ms.n[PV_MAP_INDIRECT];
        KdTree<Photon> **maps[2] = {&causticMap, &indirectMap};
        const int which[2] = {PV_MAP_CAUSTIC, PV_MAP_INDIRECT};
        for (int k = 0; k < 2; ++k) {
            uint64_t n = ms.n[which[k]], got = 0;
            delete *maps[k]; *maps[k] = NULL;
            g_pv.surf_pos[k].clear(); g_pv.surf_wi[k].clear(); g_pv.surf_alpha[k].clear();
            if (!n) continue;
            std::vector<float> pos(3 * n), wi(3 * n), alpha((size_t)PV_NSPEC * n);
            rc = pv_get_map_photons(g_pv.ctx, which[k], pos.data(), wi.data(), alpha.data(), NULL, n, &got);
            if (rc) pv_fail("pv_get_map_photons", rc);
            vector<Photon> photons(got);
            for (uint64_t i = 0; i < got; ++i) {
                Photon &p = photons[i];
                p.p = Point(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]); p.wi = Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
                p.alpha = Spectrum(0.f); memcpy(p.alpha.c, &alpha[(size_t)PV_NSPEC * i], sizeof(float) * PV_NSPEC);
            }
            *maps[k] = new KdTree<Photon>(photons);
            // kept for the LPhoton lookups of primary hits (SamplerRenderer::Render)
            g_pv.surf_pos[k].assign(pos.begin(), pos.begin() + 3 * got); g_pv.surf_wi[k].assign(wi.begin(), wi.begin() + 3 * got);
            g_pv.surf_alpha[k].assign(alpha.begin(), alpha.begin() + (size_t)PV_NSPEC * got);
        }
        delete radianceMap; radianceMap = NULL;
        uint64_t nrad = ms.n[PV_MAP_RADIANCE];
        if (finalGather && nrad) {
            // ComputeRadianceTask on the device (photonshooter.cpp:359-395, 494-524)
            std::vector<float> pos(3 * nrad), nrm(3 * nrad), Lo((size_t)PV_NSPEC * nrad);
            uint64_t got = 0;
            rc = pv_radiance_photons(g_pv.ctx, nLookup, maxDistSquared, NULL, Lo.data(), nrad, &got);
            if (rc) pv_fail("pv_radiance_photons", rc);
            rc = pv_get_map_photons(g_pv.ctx, PV_MAP_RADIANCE, pos.data(), nrm.data(), NULL, NULL, nrad, &got);
            if (rc) pv_fail("pv_get_map_photons", rc);
            vector<RadiancePhoton> rps(got);
            for (uint64_t i = 0; i < got; ++i) {
                rps[i] = RadiancePhoton(Point(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), Normal(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
                memcpy(rps[i].Lo.c, &Lo[(size_t)PV_NSPEC * i], sizeof(float) * PV_NSPEC);
            }
            radianceMap = new KdTree<RadiancePhoton>(rps);
            g_pv.rad_pos.swap(pos); g_pv.rad_nrm.swap(nrm); g_pv.rad_Lo.swap(Lo);
            g_pv.rad_pos.resize(3 * got); g_pv.rad_nrm.resize(3 * got); g_pv.rad_Lo.resize((size_t)PV_NSPEC * got);
        } else { g_pv.rad_pos.clear(); g_pv.rad_nrm.clear(); g_pv.rad_Lo.clear(); }
        rc = pv_build(g_pv.ctx, vi->maxDist, (uint32_t)vi->nUsed);
        if (rc) pv_fail("pv_build", rc);
        fprintf(stderr, "[pv] all maps on the GPU: %llu volume, %llu caustic, %llu indirect, %llu direct photons, %llu radiance photons from %llu light "
                        "paths in %.3f s (device %.3f s, %llu blocks replayed)\n",
                (unsigned long long)ms.n[0], (unsigned long long)ms.n[1], (unsigned long long)ms.n[2], (unsigned long long)ms.n[3],
                (unsigned long long)ms.n[4], (unsigned long long)ms.nshot, now_s() - t0, ms.shoot.seconds, (unsigned long long)ms.replayed_blocks);
    } else if (nVolumePhotonsWanted > 0) {
        pv_shoot_stats st;
        double t0 = now_s();
        rc = pv_shoot(g_pv.ctx, nVolumePhotonsWanted, &prm, &st);
        if (rc == PV_ENOPHOTONS) Error("Unable to store enough photons.  Giving up.\n");      // photonshooter.cpp:292
        else if (rc) pv_fail("pv_shoot", rc);
        nVolumePaths = (int)st.paths;
        rc = pv_build(g_pv.ctx, vi->maxDist, (uint32_t)vi->nUsed);
        if (rc) pv_fail("pv_build", rc);
        uint64_t n = 0; pv_photon_count(g_pv.ctx, &n);
        fprintf(stderr, "[pv] shot %llu volume photons from %llu light paths in %.3f s (device %.3f s), map built\n",
                (unsigned long long)n, (unsigned long long)st.paths, now_s() - t0, st.seconds);
    } else {
        rc = pv_set_photons(g_pv.ctx, NULL, NULL, NULL, 0);
        if (!rc) rc = pv_build(g_pv.ctx, vi->maxDist, (uint32_t)vi->nUsed);
        if (rc) pv_fail("pv_build", rc);
    }
    g_pv.ready = true;
    pv_replicate(true, vi->maxDist, (uint32_t)vi->nUsed);
    // PV_SURFACE_MAPS=cpu: the surface maps from the reference's own CPU pass instead.  That pass keeps scattering photons in
    // the medium only while its own volume map is not full (`scatter && !volumeDone`, photonshooter.cpp:96), so it is asked for
    // as many volume photons as surface photons, never for the full volume count; its volume photons are dropped.
#ifdef PV_WITH_CPU_ROUTES
    if (surface_wanted && !surface_gpu) {
        uint32_t keep = nVolumePhotonsWanted;
        nVolumePhotonsWanted = std::min(keep, std::max(nCausticPhotonsWanted, nIndirectPhotonsWanted));
        int keepPaths = nVolumePaths;
        _ZN13PhotonShooter13RefPreprocessEPK5ScenePK6CameraPK8Renderer(this, scene, camera, renderer);
        nVolumePhotonsWanted = keep; nVolumePaths = keepPaths;
        delete volumeMap; volumeMap = NULL;
    }
#endif
}

// ------------------------------------------------------------------ PhotonVolumeIntegrator (integrators/photonvolume.cpp)
void PhotonVolumeIntegrator::RequestSamples(Sampler *sampler, Sample *sample, const Scene *scene) {
    tauSampleOffset = sample->Add1D(1);
    scatterSampleOffset = sample->Add1D(1);
}

Spectrum PhotonVolumeIntegrator::Transmittance(const Scene *scene, const Renderer *renderer, const RayDifferential &ray,
                                               const Sample *sample, RNG &rng, MemoryArena &arena) const {
    if (!scene->volumeRegion) return Spectrum(1.f);
    float step = sample ? stepSize : 4.f * stepSize;
    float offset = sample ? sample->oneD[tauSampleOffset][0] : rng.RandomFloat();
    return Exp(-scene->volumeRegion->tau(ray, step, offset));
}

static void pv_fill_ray(const RayDifferential &ray, float u_scatter, pv_ray *r) {
    r->o[0] = ray.o.x; r->o[1] = ray.o.y; r->o[2] = ray.o.z;
    r->d[0] = ray.d.x; r->d[1] = ray.d.y; r->d[2] = ray.d.z;
    r->mint = ray.mint; r->maxt = ray.maxt; r->time = ray.time; r->u_scatter = u_scatter;
}

// ---- single-ray calls, gathered across the render threads --------------------------------------------------------------
// Rays spawned by specular bounces reach the volume integrator one at a time from inside the reference's recursive surface
// shading (SpecularReflect/Transmit -> Renderer::Li -> VolumeIntegrator::Li), on every render thread at once.  A device round
// trip per ray, serialised by the context's mutex, is what such a render would spend its time on (measured: 3.2 of 3.3 s).
// Instead the first thread that arrives opens a batch and waits a few tens of microseconds for the other threads' rays; the
// batch is ONE device call; everybody picks its own result up.  Every ray goes down with a stream index of its own
// (pv_gather_indexed), drawn from its render task's RNG, so what a ray sees does not depend on the batch it joined.
namespace {
struct LiBatch {
    std::vector<pv_ray> rays; std::vector<uint64_t> index;
    std::vector<float> L, T;
    bool closed, done; int rc;
    std::condition_variable cv;
    LiBatch() : closed(false), done(false), rc(0) {}
};
struct LiBatcher {
    std::mutex mu;
    std::shared_ptr<LiBatch> open;
    std::condition_variable arrived;
    unsigned long long calls, batches;
    double seconds;                      // inside pv_gather
    bool in_flight;                      // a batch is on the device: the open one keeps collecting until it is back
    LiBatcher() : calls(0), batches(0), seconds(0), in_flight(false) {}
};
LiBatcher g_li;
const size_t kLiBatchMax = 4096;
}  // namespace

static int pv_li_batched(const pv_ray &r, uint64_t index, const pv_gather_params &prm0, float *L, float *T) {
    std::unique_lock<std::mutex> lock(g_li.mu);
    g_li.calls++;
    std::shared_ptr<LiBatch> b = g_li.open;
    const bool leader = !b;
    if (leader) { b = std::make_shared<LiBatch>(); g_li.open = b; }
    const size_t slot = b->rays.size();
    b->rays.push_back(r); b->index.push_back(index);
    if (!leader) {
        if (b->rays.size() >= kLiBatchMax) { b->closed = true; g_li.open.reset(); }
        g_li.arrived.notify_all();
        b->cv.wait(lock, [&] { return b->done; });
    } else {
        // wait for company: while another batch is on the device (one warp marches one ray, so a call takes as long as its
        // longest ray -- of the order of a millisecond -- however few rays it carries), and then until every core has a ray in or
        // 100 us have passed without a new arrival
        const size_t want = (size_t)std::max(1, NumSystemCores());
        while (!b->closed) {
            if (g_li.in_flight) { g_li.arrived.wait_for(lock, std::chrono::microseconds(200)); continue; }
            if (b->rays.size() >= want) break;
            const size_t before = b->rays.size();
            g_li.arrived.wait_for(lock, std::chrono::microseconds(100));
            if (!g_li.in_flight && b->rays.size() == before) break;
        }
        if (!b->closed) { b->closed = true; g_li.open.reset(); }
        while (g_li.in_flight) g_li.arrived.wait_for(lock, std::chrono::microseconds(200));
        g_li.in_flight = true;
        g_li.batches++;
        const size_t n = b->rays.size();
        b->L.resize(n * PV_NSPEC); b->T.resize(n * PV_NSPEC);
        lock.unlock();
        // every ray of the batch keeps the stream index its own thread gave it (pv_gather_indexed): what a ray sees does not depend
        // on the batch it joined
        pv_gather_params prm = prm0; prm.ray_index_base = 0;
        const double t0 = now_s();
        int rc = g_pv.volint < 0 ? pv_gather_indexed(g_pv.ctx, b->rays.data(), b->index.data(), n, &prm, b->L.data(), b->T.data())
                                 : pv_volume_li_indexed(g_pv.ctx, g_pv.volint, b->rays.data(), b->index.data(), n, &prm, b->L.data(), b->T.data());
        const double dt = now_s() - t0;
        lock.lock();
        g_li.seconds += dt;
        g_li.in_flight = false;
        g_li.arrived.notify_all();
        b->rc = rc; b->done = true;
        b->cv.notify_all();
    }
    if (b->rc) return b->rc;
    memcpy(L, &b->L[slot * PV_NSPEC], sizeof(float) * PV_NSPEC); memcpy(T, &b->T[slot * PV_NSPEC], sizeof(float) * PV_NSPEC);
    return 0;
}

// Single-ray form (kept so every caller of the VolumeIntegrator interface still works); the renderer below batches.
static Spectrum pv_li_one(const char *who, const Scene *scene, const RayDifferential &ray, float u_scatter, float stepSize, int nUsed,
                          float maxDist, RNG &rng, Spectrum *T) {
    if (!scene->volumeRegion) { *T = 1.f; return 0.f; }
    if (!g_pv.ready) Severe("%s::Li: the GPU context was not set up (no CUDA device?)", who);
    pv_ray r; pv_fill_ray(ray, u_scatter, &r);
    pv_gather_params prm; memset(&prm, 0, sizeof(prm));
    prm.stepsize = stepSize; prm.nused = (uint32_t)nUsed; prm.maxdist = maxDist; prm.seed = g_pv.seed;
    prm.ray_index_base = ((uint64_t)rng.RandomUInt() << 20) | 0x8000000000000000ull;     // a stream of its own per call
    float L[PV_NSPEC], Tr[PV_NSPEC];
    int rc = pv_li_batched(r, prm.ray_index_base, prm, L, Tr);
    if (rc) pv_fail(g_pv.volint < 0 ? "pv_gather" : "pv_volume_li", rc);
    Spectrum Lv(0.f);
    memcpy(Lv.c, L, sizeof(L)); memcpy(T->c, Tr, sizeof(Tr));
    return Lv;
}
Spectrum PhotonVolumeIntegrator::Li(const Scene *scene, const Renderer *renderer, const RayDifferential &ray, const Sample *sample,
                                    RNG &rng, Spectrum *T, MemoryArena &arena) const {
    return pv_li_one("PhotonVolumeIntegrator", scene, ray, sample->oneD[scatterSampleOffset][0], stepSize, nUsed, maxDist, rng, T);
}

PhotonVolumeIntegrator *CreatePhotonVolumeIntegrator(const ParamSet &params, PhotonShooter *phs) {
    float stepSize = params.FindOneFloat("stepsize", 1.f);
    int nUsed = params.FindOneInt("nused", 250);
    float maxDist = params.FindOneFloat("maxdist", 0.1f);
    return new PhotonVolumeIntegrator(stepSize, nUsed, maxDist, phs);
}

// ------------------------------------------------------------------ SingleScatteringIntegrator (integrators/single.cpp),
//                                                                    EmissionIntegrator (integrators/emission.cpp)
// the reference's other two VolumeIntegrator plugins, same declarations (integrators/single.h:44-60,
// emission.h), same .pbrt parameter ("stepsize", single.cpp:141-144, emission.cpp:109-112); Li runs on the device
// (pv_volume_li), Transmittance stays the reference's three lines like PhotonVolumeIntegrator's above.
void SingleScatteringIntegrator::RequestSamples(Sampler *sampler, Sample *sample, const Scene *scene) {
    tauSampleOffset = sample->Add1D(1);
    scatterSampleOffset = sample->Add1D(1);
}
Spectrum SingleScatteringIntegrator::Transmittance(const Scene *scene, const Renderer *renderer, const RayDifferential &ray,
                                                   const Sample *sample, RNG &rng, MemoryArena &arena) const {
    if (!scene->volumeRegion) return Spectrum(1.f);
    float step = sample ? stepSize : 4.f * stepSize;
    float offset = sample ? sample->oneD[tauSampleOffset][0] : rng.RandomFloat();
    return Exp(-scene->volumeRegion->tau(ray, step, offset));
}
Spectrum SingleScatteringIntegrator::Li(const Scene *scene, const Renderer *renderer, const RayDifferential &ray, const Sample *sample,
                                        RNG &rng, Spectrum *T, MemoryArena &arena) const {
    return pv_li_one("SingleScatteringIntegrator", scene, ray, sample->oneD[scatterSampleOffset][0], stepSize, 0, 0.f, rng, T);
}
SingleScatteringIntegrator *CreateSingleScatteringIntegrator(const ParamSet &params) {
    return new SingleScatteringIntegrator(params.FindOneFloat("stepsize", 1.f));
}
void EmissionIntegrator::RequestSamples(Sampler *sampler, Sample *sample, const Scene *scene) {
    tauSampleOffset = sample->Add1D(1);
    scatterSampleOffset = sample->Add1D(1);
}
Spectrum EmissionIntegrator::Transmittance(const Scene *scene, const Renderer *renderer, const RayDifferential &ray, const Sample *sample,
                                           RNG &rng, MemoryArena &arena) const {
    if (!scene->volumeRegion) return Spectrum(1.f);
    float step = sample ? stepSize : 4.f * stepSize;
    float offset = sample ? sample->oneD[tauSampleOffset][0] : rng.RandomFloat();
    return Exp(-scene->volumeRegion->tau(ray, step, offset));
}
Spectrum EmissionIntegrator::Li(const Scene *scene, const Renderer *renderer, const RayDifferential &ray, const Sample *sample, RNG &rng,
                                Spectrum *T, MemoryArena &arena) const {
    return pv_li_one("EmissionIntegrator", scene, ray, sample->oneD[scatterSampleOffset][0], stepSize, 0, 0.f, rng, T);
}
EmissionIntegrator *CreateEmissionVolumeIntegrator(const ParamSet &params) {
    return new EmissionIntegrator(params.FindOneFloat("stepsize", 1.f));
}

// The device context of a render whose volume integrator is "single" / "emission": the scene alone (no shooter exists for such
// a scene file unless the surface integrator is a photon map, whose CPU pass is then the reference's own).
static void pv_setup_volint(const Scene *scene, const PvVolInt &vi) {
    std::string err;
    if (!pv_export_and_set(scene, err)) {
        // "emission" is pbrt's default volume integrator and reads only the medium: a scene whose surfaces or lights are off
        // this path (area lights, other shapes ...) still runs on the device, with the medium alone exported
        std::string err2;
        if (vi.kind != PV_VOLINT_EMISSION || !pv_export_and_set(scene, err2, true)) Severe("%s", err.c_str());
        Warning("%s -- the emission integrator needs the medium only, exporting that", err.c_str());
    }
    const char *seed = getenv("PV_SEED");
    g_pv.seed = seed ? strtoull(seed, NULL, 0) : 0;
    g_pv.stepsize = vi.stepSize; g_pv.volint = vi.kind; g_pv.ready = true;
    pv_replicate(false, 0.f, 0);
}

// ------------------------------------------------------------------ SamplerRenderer (renderers/samplerrenderer.cpp)
SamplerRenderer::SamplerRenderer(Sampler *s, Camera *c, SurfaceIntegrator *si, VolumeIntegrator *vi, bool visIds, PhotonShooter *ps) {
    sampler = s; camera = c; surfaceIntegrator = si; volumeIntegrator = vi; visualizeObjectIds = visIds; photonShooter = ps;
}
SamplerRenderer::~SamplerRenderer() {
    delete sampler; delete camera; delete surfaceIntegrator; delete volumeIntegrator; delete photonShooter;
}

static Spectrum pv_surface_term(const SamplerRenderer *r, const Scene *scene, const RayDifferential &ray, const Sample *sample, RNG &rng,
                                MemoryArena &arena, Intersection *isect, int fg_task = -1, uint32_t rec = 0) {
    // first half of SamplerRenderer::Li (:239-246): note scene->Intersect shrinks ray.maxt to the hit (primitive.cpp:172)
    if (scene->Intersect(ray, isect)) {
        if (fg_task >= 0) {  // a primary hit: Li without the terms that run on the device, whose rays / lookups are queued
            if (g_fg->direct) return pv_hit_li(r, scene, ray, *isect, sample, rng, arena, fg_task, rec, Spectrum(1.f), -1);
            if (g_fg->gather_rays) pv_queue_final_gather(g_fg->full, ray, *isect, sample, arena, rec, g_fg->rays[fg_task]);
            if (g_fg->lookup[0] || g_fg->lookup[1]) {
                if (!pv_queue_lookup(ray, *isect, rng, arena, rec, g_fg->lookups[fg_task]))
                    Severe("pv: a surface with glossy components reached the device LPhoton (materials other than matte / glass are off this path)");
            }
            return g_fg->primary->Li(scene, r, ray, *isect, sample, rng, arena);
        }
        return r->surfaceIntegrator->Li(scene, r, ray, *isect, sample, rng, arena);
    }
    Spectrum Li = 0.f;
    for (uint32_t i = 0; i < scene->lights.size(); ++i) Li += scene->lights[i]->Le(ray);
    return Li;
}

Spectrum SamplerRenderer::Li(const Scene *scene, const RayDifferential &ray, const Sample *sample, RNG &rng, MemoryArena &arena,
                             Intersection *isect, Spectrum *T) const {
    Spectrum localT; if (!T) T = &localT;
    Intersection localIsect; if (!isect) isect = &localIsect;
    if (tl_path.on && g_fg && g_fg->direct) {
        // A ray behind a specular bounce of a hit shaded by pv_hit_li.  Nothing is evaluated here: the ray's volume term joins the
        // group's ONE pv_gather_indexed call (PvSecondary), the hit it ends in queues its terms under the ray's number, and what
        // this function would hand back up the reference's recursion -- T * (Le + deeper bounces) + Lv -- reaches the pixel's record
        // through those queues.  (Le is zero on every surface and on every miss here: the direct mode is on only in scenes whose
        // lights are all delta lights.)
        const PvPath here = tl_path;
        const bool hit = scene->Intersect(ray, isect);                       // shrinks ray.maxt to the hit (primitive.cpp:172)
        if (!g_fg->defer) {
            // A/B route (PV_DEFER_SECONDARY=0): the ray's volume term as a blocking batched call, its transmittance folded into the
            // throughput of the next hit at once, Le and Lv handed back up the reference's recursion
            Spectrum Lvi = volumeIntegrator->Li(scene, this, ray, sample, rng, T, arena);
            Spectrum Ls(0.f);
            if (hit) Ls = pv_hit_li(this, scene, ray, *isect, sample, rng, arena, here.task, here.rec, here.W * *T, -1);
            else for (uint32_t i = 0; i < scene->lights.size(); ++i) Ls += scene->lights[i]->Le(ray);
            tl_path = here;
            return *T * Ls + Lvi;
        }
        PvSecondary sr; sr.rec = here.rec; sr.parent = here.slot; sr.W = here.W;
        pv_fill_ray(ray, sample->oneD[g_fg->scatter_offset][0], &sr.ray);
        sr.index = ((uint64_t)rng.RandomUInt() << 20) | 0x8000000000000000ull;    // a stream of its own, from the task's RNG (as pv_li_one draws it)
        std::vector<PvSecondary> &sec = g_fg->secondary[here.task];
        const int slot = (int)sec.size();
        sec.push_back(sr);
        if (hit) pv_hit_li(this, scene, ray, *isect, sample, rng, arena, here.task, here.rec, here.W, slot);
        tl_path = here;
        *T = Spectrum(1.f);
        return Spectrum(0.f);
    }
    Spectrum Li = pv_surface_term(this, scene, ray, sample, rng, arena, isect);
    Spectrum Lvi = volumeIntegrator->Li(scene, this, ray, sample, rng, T, arena);
    return *T * Li + Lvi;
}
Spectrum SamplerRenderer::Transmittance(const Scene *scene, const RayDifferential &ray, const Sample *sample, RNG &rng,
                                        MemoryArena &arena) const {
    return volumeIntegrator->Transmittance(scene, this, ray, sample, rng, arena);
}

// One image tile (the reference's task, :60-164) up to the volume term: camera ray + surface radiance per sample.
// When the GPU path is active the samples are queued (in task order, so the frame is reproducible); otherwise the task
// finishes the sample exactly like the reference.
void SamplerRendererTask::Run() {
    Sampler *sampler = mainSampler->GetSubSampler(taskNum, taskCount);
    if (!sampler) { reporter.Update(); return; }
    const SamplerRenderer *sr = static_cast<const SamplerRenderer *>(renderer);
    PvVolInt vi;
    const bool batch = g_records && pv_volint_of(sr->volumeIntegrator, &vi) && g_pv.ready && scene->volumeRegion && !visualizeObjectIds;
    MemoryArena arena;
    RNG rng(taskNum);
    int maxSamples = sampler->MaximumSampleCount();
    Sample *samples = origSample->Duplicate(maxSamples);
    RayDifferential *rays = new RayDifferential[maxSamples];
    Spectrum *Ls = new Spectrum[maxSamples];
    Spectrum *Ts = new Spectrum[maxSamples];
    Intersection *isects = new Intersection[maxSamples];
    std::vector<PvRecord> *out = batch ? &(*g_records)[taskNum] : NULL;
    int sampleCount;
    while ((sampleCount = sampler->GetMoreSamples(samples, rng)) > 0) {
        for (int i = 0; i < sampleCount; ++i) {
            float rayWeight = camera->GenerateRayDifferential(samples[i], &rays[i]);
            rays[i].ScaleDifferentials(1.f / sqrtf(sampler->samplesPerPixel));
            if (batch) {
                PvRecord rec;
                rec.imageX = samples[i].imageX; rec.imageY = samples[i].imageY; rec.rayWeight = rayWeight;
                rec.Ls = 0.f;
                if (rayWeight > 0.f)
                    rec.Ls = pv_surface_term(sr, scene, rays[i], &samples[i], rng, arena, &isects[i], g_fg ? taskNum : -1,
                                             (uint32_t)out->size());
                pv_fill_ray(rays[i], samples[i].oneD[vi.scatterSampleOffset][0], &rec.ray);
                out->push_back(rec);
                continue;
            }
            if (rayWeight > 0.f) Ls[i] = rayWeight * renderer->Li(scene, rays[i], &samples[i], rng, arena, &isects[i], &Ts[i]);
            else { Ls[i] = 0.f; Ts[i] = 1.f; }
            if (Ls[i].HasNaNs() || Ls[i].y() < -1e-5 || isinf(Ls[i].y())) Ls[i] = Spectrum(0.f);
        }
        if (!batch && sampler->ReportResults(samples, rays, Ls, isects, sampleCount))
            for (int i = 0; i < sampleCount; ++i) camera->film->AddSample(samples[i], Ls[i]);
        arena.FreeAll();
    }
    camera->film->UpdateDisplay(sampler->xPixelStart, sampler->yPixelStart, sampler->xPixelEnd + 1, sampler->yPixelEnd + 1);
    delete sampler;
    delete[] samples; delete[] rays; delete[] Ls; delete[] Ts; delete[] isects;
    reporter.Update();
}

void SamplerRenderer::Render(const Scene *scene) {
    if (photonShooter != NULL) photonShooter->Preprocess(scene, camera, this);
    surfaceIntegrator->Preprocess(scene, camera, this);
    volumeIntegrator->Preprocess(scene, camera, this);
    PvVolInt vint;
    const bool have_vint = pv_volint_of(volumeIntegrator, &vint);
    if (have_vint && vint.kind >= 0 && scene->volumeRegion && !visualizeObjectIds) pv_setup_volint(scene, vint);
    Sample *sample = new Sample(sampler, surfaceIntegrator, volumeIntegrator, scene);
    camera->AutoFocus(this, scene, sample);
    // ---- final gathering of primary hits on the GPU?  Needs the photon-volume context, a PhotonIntegrator with final gathering, an
    // indirect map with at least the 50 photons the importance lookup insists on (photonmap.cpp:209-214) and radiance photons.
    PvFinalGather fg;
    PhotonVolumeIntegrator *pvi = dynamic_cast<PhotonVolumeIntegrator *>(volumeIntegrator);
    PhotonIntegrator *pmi = dynamic_cast<PhotonIntegrator *>(surfaceIntegrator);
#ifdef PV_WITH_CPU_ROUTES
    const char *fg_mode = getenv("PV_FINAL_GATHER");                          // "cpu": keep the reference's per-ray final gather (debugging build only)
#else
    const char *fg_mode = NULL;
#endif
    const bool fg_on = g_pv.ready && pvi && pmi && pmi->finalGather && pmi->photonShooter && pmi->photonShooter->indirectMap &&
                       pmi->photonShooter->radianceMap && !g_pv.rad_pos.empty() && g_pv.n_indirect >= 50 &&
                       !(fg_mode && !strcmp(fg_mode, "cpu")) && !visualizeObjectIds;
    // ---- LPhoton of primary hits on the GPU: the caustic term (photonmap.cpp:179) whenever the device pass made a caustic map, and
    // the indirect term (:308) when final gathering is off (with it on, the indirect map only guides the gather rays).  Both maps
    // came from pv_shoot_maps, so the scene's materials are matte / glass and every lookup takes LPhoton's diffuse branch.
    bool lp_on[2] = {false, false};
    if (g_pv.ready && pvi && pmi && pmi->photonShooter && !visualizeObjectIds && !(fg_mode && !strcmp(fg_mode, "cpu"))) {
        lp_on[0] = pmi->photonShooter->causticMap && !g_pv.surf_pos[0].empty();
        lp_on[1] = !pmi->finalGather && pmi->photonShooter->indirectMap && !g_pv.surf_pos[1].empty();
    }
    // ---- direct lighting of primary hits on the GPU: when every light is a delta light (point / spot / distant) and every other
    // term of Li that costs time is on the device as well, primary hits are shaded by pv_primary_li instead of the clone
    bool direct_on = false;
    if (g_pv.ready && pvi && pmi && pmi->photonShooter && !visualizeObjectIds && !(fg_mode && !strcmp(fg_mode, "cpu")) &&
        scene->lights.size() > 0 && g_pv.scene.desc.n_nodes > 0) {
        direct_on = true;
        for (size_t i = 0; i < scene->lights.size(); ++i) direct_on = direct_on && scene->lights[i]->IsDeltaLight();
        const PhotonShooter *psh = pmi->photonShooter;
        if (psh->causticMap && !lp_on[0]) direct_on = false;
        if (psh->indirectMap && !(pmi->finalGather ? fg_on : lp_on[1])) direct_on = false;
    }
    // Render threads.  In the direct mode nothing a render thread does waits for the device (every device call is made between
    // groups of tasks), so the reference's own thread count stands.  Otherwise (a scene with an area light, or a volume integrator
    // other than photonvolume under a recursive surface integrator) a thread spends most of a specular bounce waiting for a batched
    // device call, and the batches are as large as there are threads waiting: 16 threads per core then, unless the user chose a
    // count (--ncores) or PV_THREADS says otherwise.  (The pool is created at the first EnqueueTasks, core/parallel.cpp:728-737.)
    if (g_pv.ready && scene->volumeRegion) {
        const char *th = getenv("PV_THREADS");
        if (th && atoi(th) > 0) PbrtOptions.nCores = atoi(th);
        else if (PbrtOptions.nCores == 0 && !direct_on) PbrtOptions.nCores = min(1024, 16 * NumSystemCores());
    }
    int nPixels = camera->film->xResolution * camera->film->yResolution;
    int nTasks = max(32 * NumSystemCores(), nPixels / (16 * 16));
    nTasks = RoundUpPow2(nTasks);
    std::vector<std::vector<PvRecord> > records(nTasks);
    g_records = &records;
    double t0 = now_s();
    if (fg_on || lp_on[0] || lp_on[1] || direct_on) {
        // member-wise copies made by the classes' own (implicit) copy constructors: a shooter that does not show the maps whose
        // terms run on the device, and an integrator that looks at it.  They share the maps with the originals and are never
        // destroyed (their destructors would delete those maps a second time).
        PhotonShooter *bare = new PhotonShooter(*pmi->photonShooter);
        if (fg_on || lp_on[1]) bare->indirectMap = NULL;
        if (lp_on[0]) bare->causticMap = NULL;
        fg.full = pmi;
        fg.primary = new PhotonIntegrator(*pmi);
        fg.primary->photonShooter = bare;
        fg.gather_rays = fg_on; fg.lookup[0] = lp_on[0]; fg.lookup[1] = lp_on[1]; fg.direct = direct_on;
        { const char *d = getenv("PV_DEFER_SECONDARY"); fg.defer = !(d && !strcmp(d, "0")); }
        fg.rays.resize(nTasks); fg.lookups.resize(nTasks); fg.shadows.resize(nTasks); fg.secondary.resize(nTasks);
        fg.scatter_offset = pvi->scatterSampleOffset;
        memset(&fg.vol_prm, 0, sizeof(fg.vol_prm));
        fg.vol_prm.stepsize = pvi->stepSize; fg.vol_prm.nused = (uint32_t)pvi->nUsed; fg.vol_prm.maxdist = pvi->maxDist; fg.vol_prm.seed = g_pv.seed;
        const int map_id[2] = {PV_MAP_CAUSTIC, PV_MAP_INDIRECT};
        for (int m = 0; m < 2; ++m) {
            if (!lp_on[m]) continue;
            if (!g_pv.surf_ctx[m]) {
                int rc = pv_create(&g_pv.surf_ctx[m], pv_device_list()[0]);
                if (rc) Severe("pv_create failed (%d): %s", rc, pv_last_error(NULL));
            }
            int rc = pv_set_map_photons(g_pv.surf_ctx[m], map_id[m], g_pv.surf_pos[m].data(), g_pv.surf_wi[m].data(), g_pv.surf_alpha[m].data(),
                                        g_pv.surf_pos[m].size() / 3);
            if (!rc) rc = pv_select_map(g_pv.surf_ctx[m], map_id[m], sqrtf(pmi->maxDistSquared), pmi->nLookup);
            if (rc) Severe("surface-lookup context failed (%d): %s", rc, pv_last_error(g_pv.surf_ctx[m]));
        }
    }
    if (fg_on) {
        if (!g_pv.fg_ctx) {
            int rc = pv_create(&g_pv.fg_ctx, pv_device_list()[0]);
            if (rc) Severe("pv_create failed (%d): %s", rc, pv_last_error(NULL));
        }
        const uint64_t nrad = g_pv.rad_pos.size() / 3;
        std::vector<float> rho((size_t)PV_NSPEC * nrad, 1.f);                  // rho_r plays no part in the lookups
        int rc = pv_set_scene(g_pv.fg_ctx, &g_pv.scene.desc);
        if (!rc) rc = pv_set_map_photons(g_pv.fg_ctx, PV_MAP_RADIANCE, g_pv.rad_pos.data(), g_pv.rad_nrm.data(), rho.data(), nrad);
        if (!rc) rc = pv_set_radiance_lo(g_pv.fg_ctx, g_pv.rad_Lo.data(), nrad);
        if (!rc) rc = pv_select_map(g_pv.fg_ctx, PV_MAP_RADIANCE, sqrtf(pmi->maxDistSquared), pmi->nLookup);
        if (rc) Severe("final-gather context failed (%d): %s", rc, pv_last_error(g_pv.fg_ctx));
    }
    if (fg.primary) g_fg = &fg;
    {
        ProgressReporter reporter(nTasks, "Rendering");
        // tasks run in groups so that the queued gather rays of a group (160 B each) stay within ~0.7 GB
        const double raysPerTask = g_fg ? ((fg.gather_rays ? 2.0 * pmi->gatherSamples : 0.0) + 2.0 + (fg.direct ? 1.2 * scene->lights.size() : 0.0)) * sampler->samplesPerPixel * (double)nPixels / nTasks : 0.0;
        const int group = g_fg ? max(1, min(nTasks, (int)(4.0e6 / max(raysPerTask, 1.0)))) : nTasks;
        std::vector<float> Lindir;
        std::vector<pv_ray> grays;
        for (int first = 0; first < nTasks; first += group) {
            const int last = min(nTasks, first + group);
            vector<Task *> renderTasks;
            for (int i = first; i < last; ++i)
                renderTasks.push_back(new SamplerRendererTask(scene, this, camera, reporter, sampler, sample, visualizeObjectIds,
                                                              nTasks - 1 - i, nTasks));
            EnqueueTasks(renderTasks);
            WaitForAllTasks();
            for (uint32_t i = 0; i < renderTasks.size(); ++i) delete renderTasks[i];
            if (!g_fg) continue;
            // ---- the rays behind specular bounces of this group: ONE pv_gather_indexed for their volume terms; Ls += W * P(parent) * Lv,
            // and P(j) = P(parent) * T(j) for the terms queued at the hits those rays end in (PvSecondary)
            std::vector<std::vector<Spectrum> > chainT(nTasks);
            {
                size_t nsec = 0;
                for (int t = 0; t < nTasks; ++t) nsec += fg.secondary[t].size();
                if (nsec) {
                    std::vector<pv_ray> rr(nsec); std::vector<uint64_t> ri(nsec); std::vector<float> sL(nsec * PV_NSPEC), sT(nsec * PV_NSPEC);
                    size_t q = 0;
                    for (int t = 0; t < nTasks; ++t)
                        for (size_t i = 0; i < fg.secondary[t].size(); ++i, ++q) { rr[q] = fg.secondary[t][i].ray; ri[q] = fg.secondary[t][i].index; }
                    double ts = now_s();
                    int rc = pv_gather_indexed(g_pv.ctx, rr.data(), ri.data(), nsec, &fg.vol_prm, sL.data(), sT.data());
                    if (rc) pv_fail("pv_gather_indexed", rc);
                    fg.secondary_seconds += now_s() - ts; fg.total_secondary += nsec; fg.secondary_calls++;
                    q = 0;
                    for (int t = 0; t < nTasks; ++t) {
                        std::vector<Spectrum> &P = chainT[t];
                        P.resize(fg.secondary[t].size());
                        for (size_t i = 0; i < fg.secondary[t].size(); ++i, ++q) {
                            const PvSecondary &sr = fg.secondary[t][i];
                            Spectrum Lv(0.f), Tr(1.f);
                            memcpy(Lv.c, &sL[q * PV_NSPEC], sizeof(float) * PV_NSPEC); memcpy(Tr.c, &sT[q * PV_NSPEC], sizeof(float) * PV_NSPEC);
                            const Spectrum Pp = sr.parent >= 0 ? Spectrum(P[sr.parent]) : Spectrum(1.f);     // parents precede their children in the list
                            records[t][sr.rec].Ls += sr.W * Pp * Lv;
                            P[i] = Pp * Tr;
                        }
                        std::vector<PvSecondary>().swap(fg.secondary[t]);
                    }
                }
            }
            // ---- the shadow rays of this group's direct lighting: ONE pv_occluded + ONE pv_transmittance, then Ls += C * V * Tr
            size_t nsh = 0;
            for (int t = 0; t < nTasks; ++t) nsh += fg.shadows[t].size();
            if (nsh) {
                std::vector<pv_ray> srays(nsh); std::vector<float> su(nsh), sT(nsh * PV_NSPEC); std::vector<uint8_t> occ(nsh);
                size_t q = 0;
                for (int t = 0; t < nTasks; ++t)
                    for (size_t i = 0; i < fg.shadows[t].size(); ++i, ++q) { srays[q] = fg.shadows[t][i].ray; su[q] = fg.shadows[t][i].u; }
                double ts = now_s();
                int rc = pv_occluded(g_pv.ctx, srays.data(), nsh, occ.data());
                if (!rc) rc = pv_transmittance(g_pv.ctx, srays.data(), nsh, 4.f * pvi->stepSize, su.data(), sT.data());
                if (rc) Severe("direct lighting on the device failed (%d): %s", rc, pv_last_error(g_pv.ctx));
                fg.shadow_seconds += now_s() - ts; fg.total_shadows += nsh;
                q = 0;
                for (int t = 0; t < nTasks; ++t) {
                    for (size_t i = 0; i < fg.shadows[t].size(); ++i, ++q) {
                        if (occ[q]) continue;                                                     // !visibility.Unoccluded(scene)
                        const PvShadow &sh = fg.shadows[t][i];
                        Spectrum Tr(0.f);
                        memcpy(Tr.c, &sT[q * PV_NSPEC], sizeof(float) * PV_NSPEC);
                        records[t][sh.rec].Ls += sh.tslot >= 0 ? Spectrum(sh.C * Tr * chainT[t][sh.tslot]) : Spectrum(sh.C * Tr);       // Ld += f * Li * Tr * |wi.n| / pdf
                    }
                    std::vector<PvShadow>().swap(fg.shadows[t]);
                }
            }
            // ---- the LPhoton lookups of this group: ONE pv_surface_lphoton per map, then Ls += Lr * rho_r / pi + Lt * rho_t / pi
            size_t nl = 0;
            for (int t = 0; t < nTasks; ++t) nl += fg.lookups[t].size();
            if (nl) {
                std::vector<float> pts(3 * nl), nfs(3 * nl), Lr(nl * PV_NSPEC), Lt(nl * PV_NSPEC);
                size_t q = 0;
                for (int t = 0; t < nTasks; ++t)
                    for (size_t i = 0; i < fg.lookups[t].size(); ++i, ++q) {
                        memcpy(&pts[3 * q], fg.lookups[t][i].p, 3 * sizeof(float)); memcpy(&nfs[3 * q], fg.lookups[t][i].nf, 3 * sizeof(float));
                    }
                const uint64_t paths[2] = {(uint64_t)pmi->photonShooter->nCausticPaths, (uint64_t)pmi->photonShooter->nIndirectPaths};
                double tl = now_s();
                for (int m = 0; m < 2; ++m) {
                    if (!fg.lookup[m]) continue;
                    int rc = pv_surface_lphoton(g_pv.surf_ctx[m], pts.data(), nfs.data(), nl, (uint32_t)pmi->nLookup, pmi->maxDistSquared, paths[m],
                                                Lr.data(), Lt.data());
                    if (rc) Severe("pv_surface_lphoton failed (%d): %s", rc, pv_last_error(g_pv.surf_ctx[m]));
                    q = 0;
                    for (int t = 0; t < nTasks; ++t)
                        for (size_t i = 0; i < fg.lookups[t].size(); ++i, ++q) {
                            const PvLookup &lk = fg.lookups[t][i];
                            Spectrum sr(0.f), st(0.f);
                            memcpy(sr.c, &Lr[q * PV_NSPEC], sizeof(float) * PV_NSPEC); memcpy(st.c, &Lt[q * PV_NSPEC], sizeof(float) * PV_NSPEC);
                            const Spectrum Ll = sr * lk.rr + st * lk.rt;                                  // photonmap.cpp:101-102
                            records[t][lk.rec].Ls += lk.tslot >= 0 ? Spectrum(Ll * chainT[t][lk.tslot]) : Ll;
                        }
                    fg.total_lookups += nl;
                }
                fg.lookup_seconds += now_s() - tl;
                for (int t = 0; t < nTasks; ++t) std::vector<PvLookup>().swap(fg.lookups[t]);
            }
            // ---- the gather rays of this group: ONE pv_final_gather, then Ls += sum C * Lindir
            size_t ng = 0;
            for (int t = 0; t < nTasks; ++t) ng += fg.rays[t].size();
            if (!ng) continue;
            grays.resize(ng); Lindir.resize(ng * PV_NSPEC);
            size_t k = 0;
            for (int t = 0; t < nTasks; ++t) for (size_t i = 0; i < fg.rays[t].size(); ++i) grays[k++] = fg.rays[t][i].ray;
            double tg = now_s();
            int rc = pv_final_gather(g_pv.fg_ctx, grays.data(), ng, 4.f * pvi->stepSize, g_pv.seed, fg.next_index, Lindir.data(), NULL);
            if (rc) Severe("pv_final_gather failed (%d): %s", rc, pv_last_error(g_pv.fg_ctx));
            fg.gpu_seconds += now_s() - tg; fg.next_index += ng; fg.total_rays += ng;
            k = 0;
            for (int t = 0; t < nTasks; ++t) {
                for (size_t i = 0; i < fg.rays[t].size(); ++i, ++k) {
                    const PvGatherRay &g = fg.rays[t][i];
                    Spectrum Li(0.f);
                    memcpy(Li.c, &Lindir[k * PV_NSPEC], sizeof(float) * PV_NSPEC);
                    records[t][g.rec].Ls += g.tslot >= 0 ? Spectrum(g.C * Li * chainT[t][g.tslot]) : Spectrum(g.C * Li);
                }
                std::vector<PvGatherRay>().swap(fg.rays[t]);
            }
        }
        reporter.Done();
    }
    if (g_li.calls) {
        fprintf(stderr, "[pv] volume term of %llu secondary rays (specular bounces) in %llu batched device calls, %.3f s inside them\n", g_li.calls,
                g_li.batches, g_li.seconds);
        g_li.calls = g_li.batches = 0; g_li.seconds = 0;
    }
    if (g_fg) {
        if (fg.gather_rays)
            fprintf(stderr, "[pv] final gathering of primary hits on the GPU: %llu gather rays in %.3f s\n", (unsigned long long)fg.total_rays,
                    fg.gpu_seconds);
        if (fg.total_secondary)
            fprintf(stderr, "[pv] volume term of %llu rays behind specular bounces: %llu device call(s) (one per group of render tasks), %.3f s\n",
                    (unsigned long long)fg.total_secondary, (unsigned long long)fg.secondary_calls, fg.secondary_seconds);
        if (fg.direct)
            fprintf(stderr, "[pv] direct lighting of primary hits on the GPU: %llu shadow rays (occlusion + transmittance) in %.3f s\n",
                    (unsigned long long)fg.total_shadows, fg.shadow_seconds);
        if (fg.lookup[0] || fg.lookup[1])
            fprintf(stderr, "[pv] LPhoton of primary hits on the GPU (%s%s%s map): %llu lookups in %.3f s\n", fg.lookup[0] ? "caustic" : "",
                    fg.lookup[0] && fg.lookup[1] ? " + " : "", fg.lookup[1] ? "indirect" : "", (unsigned long long)fg.total_lookups, fg.lookup_seconds);
        g_fg = NULL;
    }
    g_records = NULL;
    // ---- the volume term of the whole frame: one pv_gather
    size_t total = 0;
    for (int t = 0; t < nTasks; ++t) total += records[t].size();
    if (total) {
        std::vector<pv_ray> rays(total);
        size_t k = 0;
        for (int t = 0; t < nTasks; ++t) for (size_t i = 0; i < records[t].size(); ++i) rays[k++] = records[t][i].ray;
        std::vector<float> L(total * PV_NSPEC), T(total * PV_NSPEC);
        pv_gather_params prm; memset(&prm, 0, sizeof(prm));
        prm.stepsize = vint.stepSize; prm.nused = (uint32_t)vint.nUsed; prm.maxdist = vint.maxDist; prm.seed = g_pv.seed;
        double t1 = now_s();
        int rc = pv_volume_term_frame(rays.data(), total, &prm, L.data(), T.data());
        if (rc) pv_fail(g_pv.volint < 0 ? "pv_gather" : "pv_volume_li", rc);
        float ms = 0.f; pv_last_kernel_ms(g_pv.ctx, &ms);
        fprintf(stderr, "[pv] surface pass %.3f s on %d cores; %s of %zu camera rays %.3f s (kernel %.3f ms)\n", t1 - t0, NumSystemCores(),
                g_pv.volint < 0 ? "volume gather" : (g_pv.volint == PV_VOLINT_SINGLE ? "single-scattering volume term" : "emission volume term"),
                total, now_s() - t1, ms);
        // ---- the film: Ls[i] = rayWeight * (T * Li + Lvi) per sample (:111,:249), the NaN / negative / infinite guard (:118-133),
        // Film::AddSample.  One thread doing that for a 26 M-sample frame (darkside.pbrt: 800 x 500 x 64 spp) took half a
        // minute, so the tiles go to the reference's task pool -- grouped into blocks at least as wide as the pixel filter reaches, in
        // FOUR rounds, the blocks of one colour of a 2 x 2 checkerboard per round: two blocks of a round are a whole block apart, the
        // filter supports of their samples cannot meet, and every pixel receives its contributions in an order that depends on the
        // tile layout alone (its own block's samples in task and sample order, neighbouring blocks' by colour) -- the image is the
        // same bit for bit for any number of threads.
        std::vector<size_t> first_k(nTasks + 1, 0);
        for (int t = 0; t < nTasks; ++t) first_k[t + 1] = first_k[t] + records[t].size();
        struct FilmTile : Task {
            std::vector<int> tasks;                      // render tasks (tiles) of this block, in task order
            const std::vector<std::vector<PvRecord> > *records; const size_t *first_k; const float *L, *T; Film *film;
            void Run() {
                for (size_t j = 0; j < tasks.size(); ++j) {
                    const std::vector<PvRecord> &recs = (*records)[tasks[j]];
                    const size_t k0 = first_k[tasks[j]];
                    for (size_t i = 0; i < recs.size(); ++i) {
                        const PvRecord &rec = recs[i];
                        const size_t k = k0 + i;
                        Spectrum Lv(0.f), Tr(1.f), Lo(0.f);
                        if (rec.rayWeight > 0.f) {
                            memcpy(Lv.c, &L[k * PV_NSPEC], sizeof(float) * PV_NSPEC); memcpy(Tr.c, &T[k * PV_NSPEC], sizeof(float) * PV_NSPEC);
                            Lo = rec.rayWeight * (Tr * rec.Ls + Lv);                  // Ls[i] = rayWeight * (T * Li + Lvi), :111,:249
                        }
                        if (Lo.HasNaNs() || Lo.y() < -1e-5 || isinf(Lo.y())) Lo = Spectrum(0.f);      // :118-133
                        CameraSample cs; cs.imageX = rec.imageX; cs.imageY = rec.imageY; cs.lensU = cs.lensV = 0.f; cs.time = rec.ray.time;
                        film->AddSample(cs, Lo);
                    }
                }
            }
        };
        // tile layout of the tasks (Sampler::ComputeSubWindow, core/sampler.cpp:55-74): nx tiles per row, read off the windows themselves
        int nx = nTasks, min_w = 1 << 30, min_h = 1 << 30;
        {
            int x0, x1, y0, y1, fy0 = 0;
            for (int t = 0; t < nTasks; ++t) {
                sampler->ComputeSubWindow(t, nTasks, &x0, &x1, &y0, &y1);
                if (t == 0) fy0 = y0;
                else if (nx == nTasks && y0 != fy0) nx = t;
                if (x1 > x0 && y1 > y0) { min_w = min(min_w, x1 - x0); min_h = min(min_h, y1 - y0); }
            }
        }
        const ImageFilm *ifilm = dynamic_cast<const ImageFilm *>(camera->film);
        const double tf = now_s();
        if (ifilm && nTasks % nx == 0 && min_w < (1 << 30)) {
            // blocks of gx x gy tiles, each at least as wide as the filter reaches across it; a block is added by one thread, tile by tile
            const int reach = 2 * Ceil2Int(max(ifilm->filter->xWidth, ifilm->filter->yWidth)) + 1;
            const int gx = (reach + min_w - 1) / min_w, gy = (reach + min_h - 1) / min_h;
            const int ny = nTasks / nx, bx = (nx + gx - 1) / gx, by = (ny + gy - 1) / gy;
            for (int colour = 0; colour < 4; ++colour) {
                vector<Task *> blocks;
                for (int b = 0; b < bx * by; ++b) {
                    const int cx = b % bx, cy = b / bx;
                    if (((cx & 1) | ((cy & 1) << 1)) != colour) continue;
                    FilmTile *ft = new FilmTile; ft->records = &records; ft->first_k = first_k.data(); ft->L = L.data(); ft->T = T.data(); ft->film = camera->film;
                    for (int ty = cy * gy; ty < min(ny, (cy + 1) * gy); ++ty)
                        for (int tx = cx * gx; tx < min(nx, (cx + 1) * gx); ++tx)
                            if (!records[ty * nx + tx].empty()) ft->tasks.push_back(ty * nx + tx);
                    if (ft->tasks.empty()) delete ft; else blocks.push_back(ft);
                }
                if (blocks.empty()) continue;
                EnqueueTasks(blocks);
                WaitForAllTasks();
                for (size_t i = 0; i < blocks.size(); ++i) delete blocks[i];
            }
        } else {
            FilmTile ft; ft.records = &records; ft.first_k = first_k.data(); ft.L = L.data(); ft.T = T.data(); ft.film = camera->film;
            for (int t = 0; t < nTasks; ++t) ft.tasks.push_back(t);                 // another film class: in task order, one thread
            ft.Run();
        }
        if (total > 4000000) fprintf(stderr, "[pv] film: %zu samples added in %.3f s\n", total, now_s() - tf);
    }
    delete sample;
    camera->film->WriteImage();
}
