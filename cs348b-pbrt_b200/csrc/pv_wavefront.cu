// pv_wavefront.cu -- photon shooting as a WAVEFRONT (PhotonShootingTask::Run / followPhoton, core/photonshooter.cpp:47-357).
//
// The light paths in flight live in P slots of SoA state in global memory (continuation frame, Philox position, frame stack);
// one generation of the wavefront is three small kernels, each of which all warps run in the same state:
//
//   wf_event_kernel   slot-parallel.  What followPhoton does once the free-flight march of a segment has ended: the medium
//                     interaction (:82-128: scatter test Q1, volume deposit, new direction, the frame push of Q2) or the surface
//                     part (:131-227: transmittance up to the hit, surface deposits, matte / glass bounce, dispersion split),
//                     and the pops of the continuation stack.  Leaves the slot with a ray to trace, or free.
//   wf_trace_kernel   slot-parallel.  Free slots draw the next light path (warp-aggregated counter) and emit it (:248-275:
//                     PermutedHalton in fp64, light choice, Sample_L); every slot with a ray intersects it with the scene
//                     (BVH walk, hit record) and clips it against the medium; segments that cross the medium are queued.
//   wf_march_kernel   persistent warps over the queue of segments: the free-flight loop of :66-80 (the reference re-marches
//                     the optical depth from the segment start with a fresh offset at every step, Q3).  A lane whose segment
//                     has ended pulls the next one from the queue at once, so the warp stays full whatever the lengths of the
//                     segments; the ray-only part of DensityRegion::tau (renormalisation, the slab distances) is computed
//                     once per segment and kept in registers.
//
// Every draw a path makes comes from its own Philox stream in the reference's depth-first order, deposits carry the id
// (path << 16 | ordinal) and are ordered by it afterwards (shoot_finish): the photon set is bit-identical to the one the
// persistent-thread kernel of pv_shoot.cu produces, whatever P and the scheduling are (tests compare both with the oracle).
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "pv_shoot.cuh"

enum { WS_NEW = 0, WS_TRACE = 1, WS_MARCHING = 2, WS_HIT = 3, WS_MISS = 4, WS_POP = 5, WS_IDLE = 6 };
#define WF_F4 13                       // float4s of a frame: o|mint, d|maxt, ip|ieps, inn|prim, idpdu|(nI, spec, loop_i), 8 x alpha
// Where float4 k of slot `slot` lives in a frame array of P slots.  Frames are RECORDS (AoS, 14 float4 = 224 B = seven 32-byte
// sectors each): the event kernel reads and writes whole frames of slots in queue order, i.e. at random, and a 16-byte access into
// a plane per component (SoA) moved a 32-byte sector for every 16 bytes used -- 5.2 GB of DRAM traffic per generation of 4 M slots.
// The slot-parallel kernels read o|mint + d|maxt (one sector) and write ip..idpdu (two and a half).
#ifndef WF_FRAME_SOA
#define WF_FRAME_SOA 0
#endif
#define WF_STRIDE 14
#if WF_FRAME_SOA
#define WF_AT(P, slot, k) ((size_t)(k) * (size_t)(P) + (slot))
#else
#define WF_AT(P, slot, k) ((size_t)(slot) * WF_STRIDE + (k))
#endif
#define WF_FRAME_F4S(P) ((size_t)WF_STRIDE * (size_t)(P))      // float4s of a frame array of P slots
#define WF_THREADS 128
#define WF_FAST_LEVELS 4                // stack levels every slot owns; deeper ones (a path scattering again and again) come out of a page pool
#ifndef WF_EVENT_MIN_CTAS
#define WF_EVENT_MIN_CTAS 6
#endif
#ifndef WF_MARCH_THREADS
#define WF_MARCH_THREADS 128
#endif
#ifndef WF_MARCH_MIN_CTAS
#define WF_MARCH_MIN_CTAS 6                // 78 registers, no spills (8: 64 registers, 36 bytes spilled; measured 97 -> 90 ms together with the event kernel's 6)
#endif

struct WaveState {
    uint32_t P;
    uint32_t volume_only;              // the surface maps are off: a diffuse bounce ends the path (Q6), nothing but volume photons is stored
    float4 *frame;                     // P frames, float4 k of slot s at WF_AT(P, s, k)
    float4 *stack;                     // [WF_FAST_LEVELS] frame arrays: the first levels of every slot's continuation stack
    float4 *deep;                      // [deep_pages][SH_MAXDEPTH - WF_FAST_LEVELS][WF_F4]: the levels above, one page per slot that ever needs them
    uint32_t *deep_page;               // [P] page of the slot (~0 = none yet); pages are handed out by a bump counter and kept for the whole wave
    uint32_t deep_pages;
    uint32_t *state;                   // [P] WS_*
    uint64_t *path;                    // [P] light-path index (1-based, global)
    uint4 *rng_buf; uint32_t *rng_jp;  // [P] Philox block in use, its successor's index j | pos << 28
    uint2 *misc;                       // [P] deposit ordinal, stack height
    float4 *march;                     // [P] t0, t1, t_i, xi of the segment being marched
    uint32_t *queue;                   // [P] slots whose segment crosses the medium
    uint32_t *equeue;                  // [2][P] slots wf_event_kernel has work for; filled by generation gen into half gen & 1
    uint32_t gen;                      // generation being launched
    unsigned int *ctr;                 // [0] queue length, [1] queue head, [2] slots not idle after wf_trace_kernel, [4 + h] length of equeue half h, [6] deep pages handed out, [7] a slot found the deep pool empty (the wave is replayed with a larger one)
};

__device__ __forceinline__ uint32_t pack_meta(int nI, int spec, int loop_i) { return (uint32_t)nI | ((uint32_t)spec << 20) | ((uint32_t)(loop_i + 1) << 24); }
__device__ __forceinline__ void frame_store(float4 *base, uint32_t P, uint32_t slot, const Frame &f) {
    base[WF_AT(P, slot, 0)] = make_float4(f.o[0], f.o[1], f.o[2], f.mint);
    base[WF_AT(P, slot, 1)] = make_float4(f.d[0], f.d[1], f.d[2], f.maxt);
    base[WF_AT(P, slot, 2)] = make_float4(f.ip[0], f.ip[1], f.ip[2], f.ieps);
    base[WF_AT(P, slot, 3)] = make_float4(f.inn[0], f.inn[1], f.inn[2], __int_as_float(f.prim));
    base[WF_AT(P, slot, 4)] = make_float4(f.idpdu[0], f.idpdu[1], f.idpdu[2], __uint_as_float(pack_meta(f.nI, f.spec, f.loop_i)));
#pragma unroll 1
    for (int q = 0; q < 7; ++q) base[WF_AT(P, slot, 5 + q)] = make_float4(f.alpha[4 * q], f.alpha[4 * q + 1], f.alpha[4 * q + 2], f.alpha[4 * q + 3]);
    base[WF_AT(P, slot, 12)] = make_float4(f.alpha[28], f.alpha[29], 0.f, 0.f);
}
// spec bit 2 of a slot's CURRENT frame (never of a stacked one): alpha is still the emission weight Le * |cos| / (pdf * lightPdf)
// (photonshooter.cpp:262-264), which the slot does not store -- float4 12 holds {|cos|, pdf * lightPdf, spot falloff, light number}
// and wf_event_kernel re-makes the 30 bins from the light's spectrum when it first needs them (most paths end without)
#define WF_SPEC_PRISTINE 4
__device__ __forceinline__ void frame_load(const float4 *base, uint32_t P, uint32_t slot, Frame &f, const DevScene *sc = nullptr) {
    float4 v = base[WF_AT(P, slot, 0)]; f.o[0] = v.x; f.o[1] = v.y; f.o[2] = v.z; f.mint = v.w;
    v = base[WF_AT(P, slot, 1)]; f.d[0] = v.x; f.d[1] = v.y; f.d[2] = v.z; f.maxt = v.w;
    v = base[WF_AT(P, slot, 2)]; f.ip[0] = v.x; f.ip[1] = v.y; f.ip[2] = v.z; f.ieps = v.w;
    v = base[WF_AT(P, slot, 3)]; f.inn[0] = v.x; f.inn[1] = v.y; f.inn[2] = v.z; f.prim = __float_as_int(v.w);
    v = base[WF_AT(P, slot, 4)]; f.idpdu[0] = v.x; f.idpdu[1] = v.y; f.idpdu[2] = v.z;
    const uint32_t m = __float_as_uint(v.w);
    f.nI = (int)(m & 0xfffffu); f.spec = (int)((m >> 20) & 0xfu); f.loop_i = (int)(m >> 24) - 1;
    if (sc && (f.spec & WF_SPEC_PRISTINE)) {
        v = base[WF_AT(P, slot, 12)];
        const pv_light &l = sc->lights[__float_as_uint(v.w)];
        const float ad = v.x, den = v.y, scale = v.z;
        const bool spot = l.type == PV_LIGHT_SPOT;
#pragma unroll 1
        for (int b = 0; b < PV_NSPEC; ++b) f.alpha[b] = __fdiv_rn((spot ? l.intensity[b] * scale : l.intensity[b]) * ad, den);
        f.spec &= ~WF_SPEC_PRISTINE;
        return;
    }
#pragma unroll 1
    for (int q = 0; q < 7; ++q) { v = base[WF_AT(P, slot, 5 + q)]; f.alpha[4 * q] = v.x; f.alpha[4 * q + 1] = v.y; f.alpha[4 * q + 2] = v.z; f.alpha[4 * q + 3] = v.w; }
    v = base[WF_AT(P, slot, 12)]; f.alpha[28] = v.x; f.alpha[29] = v.y;
}
__device__ __forceinline__ void rng_store(const WaveState &w, uint32_t slot, const PathRng &r) {
    w.rng_buf[slot] = make_uint4(r.buf[0], r.buf[1], r.buf[2], r.buf[3]);
    w.rng_jp[slot] = r.j | (r.pos << 28);
}
__device__ __forceinline__ void rng_load(const WaveState &w, uint32_t slot, uint64_t path, uint32_t k0, uint32_t k1, PathRng &r) {
    r.c0 = (uint32_t)path; r.c1 = (uint32_t)(path >> 32); r.k0 = k0; r.k1 = k1;
    const uint4 b = w.rng_buf[slot]; r.buf[0] = b.x; r.buf[1] = b.y; r.buf[2] = b.z; r.buf[3] = b.w;
    const uint32_t jp = w.rng_jp[slot]; r.j = jp & 0x0fffffffu; r.pos = jp >> 28;
}
__device__ __forceinline__ void flush_stats(unsigned long long *stats, const uint32_t *c, int n) {
    for (int i = 0; i < n; ++i) {
        unsigned long long v = c[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(PV_FULL, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&stats[i], v);
    }
}

// a slot whose march or trace has ended in something wf_event_kernel must handle
__device__ __forceinline__ void event_push(const WaveState &w, uint32_t slot) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned int qi = 0;
    if (g.thread_rank() == 0) qi = atomicAdd(&w.ctr[4 + (w.gen & 1u)], (unsigned int)g.size());
    qi = g.shfl(qi, 0) + g.thread_rank();
    w.equeue[(size_t)(w.gen & 1u) * w.P + qi] = slot;
}

// ---------------------------------------------------------------------------------------------------------------- trace
// KIND: bit 0 = the scene holds sphere primitives, bit 1 = the medium is exponential (as in pv_shoot.cu)
template <int KIND>
__global__ void __launch_bounds__(WF_THREADS) wf_trace_kernel(ShootArgs a, WaveState w) {
    constexpr bool SPH = (KIND & 1) != 0;
    typedef typename std::conditional<(KIND & 2) != 0, MedView, MedViewPlain>::type MV;
    __shared__ uint32_t s_perm[41];
    if (threadIdx.x < 41) s_perm[threadIdx.x] = a.perm[threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) w.ctr[4 + ((w.gen + 1u) & 1u)] = 0;      // the half wf_event_kernel has just consumed: next generation's
    __syncthreads();
    const DevScene &sc = *a.sc;
    const MV med = make_medview<MV>(sc.med);
    const uint32_t slot = blockIdx.x * WF_THREADS + threadIdx.x;
    const bool in = slot < w.P;
    uint32_t st = in ? w.state[slot] : (uint32_t)WS_IDLE;
    uint32_t cnt[6] = {0, 0, 0, 0, 0, 0};          // nodes, tris, density samples, segments, overflows, paths
    const uint64_t total = (uint64_t)a.n_local_blocks * SH_BLOCK;
    PathRng rng; rng.reset(0, a.k0, a.k1);
    v3 o = V3(0.f, 0.f, 0.f), d = V3(0.f, 0.f, 1.f);
    float mint = 0.f, maxt = INFINITY;
    int nI = 0, spec = 0;
    uint32_t sp = 0;
    uint64_t path = 0;
    float4 em = make_float4(0.f, 0.f, 0.f, 0.f);
    bool live = false, fresh = false;
    if (st == WS_NEW) {
        // ---- fetch a light path (warp-aggregated counter) and emit it: photonshooter.cpp:248-275
        cg::coalesced_group g = cg::coalesced_threads();
        unsigned long long wk = 0;
        if (g.thread_rank() == 0) wk = atomicAdd(a.work, (unsigned long long)g.size());
        wk = g.shfl(wk, 0) + g.thread_rank();
        if (wk >= total) { st = WS_IDLE; w.state[slot] = WS_IDLE; }
        else {
            const uint32_t lblock = (uint32_t)(wk / SH_BLOCK);
            const uint64_t gblock = a.b_start + (uint64_t)lblock * a.world;
            path = (gblock - 1) * SH_BLOCK + (wk % SH_BLOCK) + 1;
            rng.reset(path, a.k0, a.k1);
            cnt[5]++;
            float u[6];
            {
                const uint32_t halton_base[6] = {2, 3, 5, 7, 11, 13};
                const uint32_t *p = s_perm;
#pragma unroll
                for (int dmn = 0; dmn < 6; ++dmn) {
                    uint32_t base = halton_base[dmn], n = (uint32_t)path;
                    double val = 0, invBase = 1. / base, invBi = invBase;
                    while (n > 0) {
                        uint32_t d_i = p[n % base];
                        val += d_i * invBi;
                        n = __double2uint_rz((double)n * invBase);
                        invBi *= invBase;
                    }
                    u[dmn] = fminf((float)val, PV_ONE_MINUS_EPS);
                    p += base;
                }
            }
            // SampleDiscrete (montecarlo.h:99-107): upper_bound on the CDF
            int nl = (int)sc.n_lights, lo = 0, hi = nl + 1;
            while (lo < hi) { int mid = (lo + hi) / 2; if (u[0] < sc.light_cdf[mid]) hi = mid; else lo = mid + 1; }
            int lightNum = max(lo - 1, 0);
            float lightPdf = __fdiv_rn(sc.light_func[lightNum], sc.light_func_int * nl);
            const pv_light &l = sc.lights[lightNum];
            v3 ro, rd; float pdf, scale = 1.f;
            v3 lnrm = V3(0.f, 0.f, 0.f); bool own_normal = false, facing = true;
            if (l.type == PV_LIGHT_AREA) {                                 // lights/diffuse.cpp:89-100, ShapeSet::Sample(ls, Ns) core/light.cpp:161-164
                ro = area_sample_point(sc, l, lightNum, u[3], u[1], u[2], &lnrm);
                rd = uniform_sample_sphere(u[4], u[5]);
                if (vdot(rd, lnrm) < 0.f) rd = rd * -1.f;
                // ShapeSet::Pdf(Point) (:175-180) sums areas[i] * (1 / areas[i]) over the shapes: the NUMBER of shapes over the total area
                pdf = __fdiv_rn(sc.larea_pd[lightNum], sc.larea_sum[lightNum]) * 0.15915494309189533577f;
                own_normal = true; facing = vdot(lnrm, rd) > 0.f;           // DiffuseAreaLight::L: Lemit towards the normal's side only
                mint = 1e-3f;                                              // Ray(org, dir, 1e-3f, INFINITY, time)
            } else if (l.type == PV_LIGHT_POINT) {                         // lights/point.cpp:80-88
                ro = V3(l.pos[0], l.pos[1], l.pos[2]); rd = uniform_sample_sphere(u[1], u[2]);
                pdf = __fdiv_rn(1.f, 4.f * PV_PI_F);
            } else if (l.type == PV_LIGHT_SPOT) {                          // lights/spot.cpp:106-114
                v3 v = uniform_sample_cone(u[1], u[2], l.cos_total_width);
                ro = V3(l.pos[0], l.pos[1], l.pos[2]); rd = xf_vec(l.light_to_world, v);
                pdf = __fdiv_rn(1.f, 2.f * PV_PI_F * (1.f - l.cos_total_width));
                scale = spot_falloff(l, rd);
            } else {                                                       // lights/distant.cpp:82-102
                const float *wb = sc.world_bound;
                v3 pmin = V3(wb[0], wb[1], wb[2]), pmax = V3(wb[3], wb[4], wb[5]);
                v3 wc = pmin * .5f + pmax * .5f;
                float wr = bbox_inside(wb, wb + 3, wc) ? vlen(wc - pmax) : 0.f;
                v3 ld = V3(l.dir[0], l.dir[1], l.dir[2]), v1, v2;
                coordinate_system(ld, &v1, &v2);
                float d1, d2;
                concentric_sample_disk(u[1], u[2], &d1, &d2);
                v3 Pdisk = wc + (v1 * d1 + v2 * d2) * wr;
                ro = Pdisk + ld * wr; rd = -ld;
                pdf = __fdiv_rn(1.f, PV_PI_F * wr * wr);
            }
            const float ad = fabsf(vdot(own_normal ? lnrm : rd, rd));        // AbsDot(Nl, photonRay.d); Nl == ray.d for the delta lights
            const float den = pdf * lightPdf;
            bool black = true;
            int npos = 0;
#pragma unroll 1
            for (int b = 0; b < PV_NSPEC; ++b) {                            // alpha = Le * |cos| / (pdf * lightPdf): tested here, re-made when first needed
                const float Le = l.type == PV_LIGHT_SPOT ? l.intensity[b] * scale : l.intensity[b];
                const float al = __fdiv_rn(Le * ad, den);
                black = black && (al == 0.f);
                npos += al > 0.f ? 1 : 0;
            }
            if (!(pdf == 0.f || black || !facing)) {                       // else: the slot stays free and draws the next path
                o = ro; d = rd; if (!own_normal) mint = 0.f; maxt = INFINITY; nI = 0; spec = 1 | (npos == 1 ? 2 : 0) | WF_SPEC_PRISTINE;
                em = make_float4(ad, den, scale, __uint_as_float((uint32_t)lightNum));
                w.path[slot] = path;
                w.misc[slot] = make_uint2(0u, 0u);
                live = true; fresh = true;
            }
        }
    } else if (st == WS_TRACE) {
        const float4 f0 = w.frame[WF_AT(w.P, slot, 0)], f1 = w.frame[WF_AT(w.P, slot, 1)];
        o = V3(f0.x, f0.y, f0.z); mint = f0.w; d = V3(f1.x, f1.y, f1.z); maxt = f1.w;
        const uint32_t m = __float_as_uint(w.frame[WF_AT(w.P, slot, 4)].w);
        nI = (int)(m & 0xfffffu); spec = (int)((m >> 20) & 0xfu);
        path = w.path[slot];
        sp = w.misc[slot].y;
        rng_load(w, slot, path, a.k0, a.k1, rng);
        live = true;
    }
    if (live) {
        // ---- followPhoton head: intersect, clip against the medium (photonshooter.cpp:54-65)
        cnt[3]++;
        float thit = maxt;
        BvhCounters bc = {0, 0};
        const int prim = bvh_traverse<false, SPH>(sc, o, d, mint, &thit, &bc);
        cnt[0] += bc.nodes; cnt[1] += bc.tris;
        uint32_t ns = WS_POP;
        if (prim >= 0) {
            // hit record: shapes/trianglemesh.cpp:160-205 with default uvs, core/diffgeom.cpp:40-55
            v3 dpdu, nn, hp; float eps;
            const float *tv = sc.tri + 9 * (size_t)prim;
            if (!SPH || tv[0] == tv[0]) {
                v3 p1 = V3(tv[0], tv[1], tv[2]), p2 = V3(tv[3], tv[4], tv[5]), p3 = V3(tv[6], tv[7], tv[8]);
                v3 dp1 = p1 - p3, dp2 = p2 - p3;
                dpdu = (dp1 * -1.f - dp2 * -1.f) * 1.f;                     // (dv2*dp1 - dv1*dp2) * invdet, dv1 = dv2 = -1
                v3 dpdv = (dp1 * -0.f + dp2 * -1.f) * 1.f;                  // (-du2*dp1 + du1*dp2) * invdet, du2 = 0, du1 = -1
                nn = vnorm(vcross(dpdu, dpdv));
                hp = ray_at(o, d, thit);
                eps = 1e-3f * thit;
            } else sphere_dg(sc.spheres + (__float_as_uint(tv[0]) & PV_SPHERE_INDEX_MASK), o, d, thit, &hp, &nn, &dpdu, &eps);   // NaN-tagged slot: shapes/sphere.cpp:112-163
            maxt = thit;                                                    // GeometricPrimitive::Intersect: r.maxt = thit
            nI++;
            // With the surface maps off, a path whose continuation stack is empty and that reaches a matte surface is over: the
            // bounce is diffuse, so `indirectDone && !specularPath` (photonshooter.cpp:218, Q6) ends it whatever is drawn, nothing
            // is deposited on surfaces and no frame is left to go on drawing from its stream.  Its transmittance up to the hit and
            // its weight are never read: the slot is free as soon as the march (if any) finds no interaction.
            const bool dies = w.volume_only && sp == 0 && sc.mats[sc.prim_mat[prim]].type == PV_MAT_MATTE;
            const float length = vlen(d);
            if (length != 0.f) {
                const v3 rnd = vdiv(d, length);
                float t0, t1;
                if (!med_intersectp(med, o, rnd, mint * length, maxt * length, &t0, &t1)) { t0 = 1.0f; t1 = 0.0f; }
                t0 += rng.next() * a.stepsize;
                const float xi = rng.next();
                ns = t0 < t1 ? WS_MARCHING : (dies ? WS_NEW : WS_MISS);
                if (ns == WS_MARCHING) {
                    w.march[slot] = make_float4(t0, t1, t0, xi);
                    cg::coalesced_group g = cg::coalesced_threads();
                    unsigned int qi = 0;
                    if (g.thread_rank() == 0) qi = atomicAdd(&w.ctr[0], (unsigned int)g.size());
                    qi = g.shfl(qi, 0) + g.thread_rank();
                    w.queue[qi] = slot | (dies ? 0x80000000u : 0u);
                }
            } else if (dies) ns = WS_NEW;
            if (ns != WS_NEW) {
                w.frame[WF_AT(w.P, slot, 2)] = make_float4(hp.x, hp.y, hp.z, eps);
                w.frame[WF_AT(w.P, slot, 3)] = make_float4(nn.x, nn.y, nn.z, __int_as_float(prim));
                w.frame[WF_AT(w.P, slot, 4)] = make_float4(dpdu.x, dpdu.y, dpdu.z, __uint_as_float(pack_meta(nI, spec, -1)));
            }
        } else if (sp == 0) ns = WS_NEW;                                    // left the scene, nothing to resume: the path is over
        if (ns != WS_NEW) {
            if (fresh || prim >= 0) {
                w.frame[WF_AT(w.P, slot, 0)] = make_float4(o.x, o.y, o.z, mint);
                w.frame[WF_AT(w.P, slot, 1)] = make_float4(d.x, d.y, d.z, maxt);
            }
            if (fresh) w.frame[WF_AT(w.P, slot, 12)] = em;
            rng_store(w, slot, rng);
            if (ns != WS_MARCHING) event_push(w, slot);
        }
        w.state[slot] = ns;
        st = ns;
    }
    const uint32_t act = __ballot_sync(PV_FULL, in && st != WS_IDLE);
    if ((threadIdx.x & 31) == 0 && act) atomicAdd(&w.ctr[2], (unsigned int)__popc(act));
    flush_stats(a.stats, cnt, 6);
}

// ---------------------------------------------------------------------------------------------------------------- march
// VolumeGridDensity::Density (volumes/volumegrid.cpp:39-57) of a grid whose extent is a power of two along every axis:
// the three divisions by the extent are multiplications by its (exact) reciprocal -- the same quotients bit for bit.
template <class Med>
__device__ __forceinline__ float grid_density_pow2(const Med &m, v3 Pobj, float ix, float iy, float iz) {
    if (!bbox_inside(m.p0, m.p1, Pobj)) return 0.f;
    float vx_ = (Pobj.x - m.p0[0]) * ix, vy_ = (Pobj.y - m.p0[1]) * iy, vz_ = (Pobj.z - m.p0[2]) * iz;
    vx_ = vx_ * m.nx - .5f; vy_ = vy_ * m.ny - .5f; vz_ = vz_ * m.nz - .5f;
    const int vx = (int)floorf(vx_), vy = (int)floorf(vy_), vz = (int)floorf(vz_);
    const float dx = vx_ - vx, dy = vy_ - vy, dz = vz_ - vz;
    const int x0 = min(max(vx, 0), m.nx - 1), x1 = min(max(vx + 1, 0), m.nx - 1);
    const int y0 = min(max(vy, 0), m.ny - 1) * m.nx, y1 = min(max(vy + 1, 0), m.ny - 1) * m.nx;
    const int sl = m.nx * m.ny;
    const float *z0 = m.density + (size_t)(min(max(vz, 0), m.nz - 1) * sl), *z1 = m.density + (size_t)(min(max(vz + 1, 0), m.nz - 1) * sl);
    const float d00 = lerpf(dx, __ldg(z0 + (y0 + x0)), __ldg(z0 + (y0 + x1)));
    const float d10 = lerpf(dx, __ldg(z0 + (y1 + x0)), __ldg(z0 + (y1 + x1)));
    const float d01 = lerpf(dx, __ldg(z1 + (y0 + x0)), __ldg(z1 + (y0 + x1)));
    const float d11 = lerpf(dx, __ldg(z1 + (y1 + x0)), __ldg(z1 + (y1 + x1)));
    return lerpf(dz, lerpf(dy, d00, d10), lerpf(dy, d01, d11));
}
__device__ __forceinline__ bool is_pow2f(float x) { return x > 0.f && (__float_as_uint(x) & 0x007fffffu) == 0u && x >= 1.1754944e-38f && x < INFINITY; }

// The free-flight loop (photonshooter.cpp:66-80).  Step k of a segment is Transmittance(Ray(p, wi, tInit, t0 + k * stepsize)) with
// a fresh offset, i.e. about (t0 - tInit + k * stepsize) / (4 * integrator step) density samples summed in order
// (DensityRegion::tau, core/volume.cpp:296-310), then the luminance test against xi (:74-77).  A segment is a few steps of a
// few samples each, so a lane alternates between four kinds of work of very different cost: PULL (finish a segment, take the
// next one from the queue and set its ray up), STEP (test of the step just sampled; offset and interval of the next one), SAMPLE
// (one trilinear density tap) and REFILL (one Philox block = four draws).  Every pass of the warp VOTES: the kind of work most
// lanes are waiting for is executed, by exactly those lanes, the others keep their state -- each pass runs one straight piece
// of code with as many lanes as any schedule could give it, instead of every lane dragging the whole warp through its own
// sequence.
enum { MS_PULL = 0, MS_STEP = 1, MS_SAMPLE = 2, MS_REFILL = 3, MS_DONE = 4 };
// the vote: the kind with the largest (waiting lanes x weight) runs; WF_SAMPLES_PER_PASS density taps per SAMPLE pass
#ifndef WF_W_SAMPLE
#define WF_W_SAMPLE 4
#endif
#ifndef WF_W_STEP
#define WF_W_STEP 4
#endif
#ifndef WF_W_PULL
#define WF_W_PULL 4
#endif
#ifndef WF_W_REFILL
#define WF_W_REFILL 4
#endif
#ifndef WF_SAMPLES_PER_PASS
#define WF_SAMPLES_PER_PASS 2
#endif
template <int KIND>
__global__ void __launch_bounds__(WF_MARCH_THREADS, WF_MARCH_MIN_CTAS) wf_march_kernel(ShootArgs a, WaveState w) {
    typedef typename std::conditional<(KIND & 2) != 0, MedView, MedViewPlain>::type MV;
    __shared__ float s_cie[PV_NSPEC], s_st[PV_NSPEC];
    __shared__ float s_minmax[3];
    const DevScene &sc = *a.sc;
    const DevMedium &gmed = sc.med;
    const MV med = make_medview<MV>(gmed);
    if (threadIdx.x < PV_NSPEC) { s_cie[threadIdx.x] = sc.cie_y[threadIdx.x]; s_st[threadIdx.x] = gmed.sigma_a[threadIdx.x] + gmed.sigma_s[threadIdx.x]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mn = INFINITY, mx = 0.f, y1 = 0.f;
#pragma unroll 1
        for (int b = 0; b < PV_NSPEC; ++b) { mn = fminf(mn, s_st[b]); mx = fmaxf(mx, s_st[b]); y1 += s_cie[b]; }
        s_minmax[0] = mn; s_minmax[1] = mx; s_minmax[2] = __fdiv_rn(y1 * 300.f, 106.856895f * (float)PV_NSPEC);
    }
    __syncthreads();
    const float st_min = s_minmax[0], st_max = s_minmax[1], y_one = s_minmax[2];
    const unsigned int n_q = w.ctr[0];
    const bool homog = med_is_homog(med);
    const float ex = med.p1[0] - med.p0[0], ey = med.p1[1] - med.p0[1], ez = med.p1[2] - med.p0[2];
    const bool pow2 = med.type == PV_MEDIUM_GRID && is_pow2f(ex) && is_pow2f(ey) && is_pow2f(ez);
    const float ix = pow2 ? 1.f / ex : 0.f, iy = pow2 ? 1.f / ey : 0.f, iz = pow2 ? 1.f / ez : 0.f;
    const float istep = a.istep4;
    uint32_t c_dens = 0;
    uint32_t ms = MS_PULL;
    bool have = false, dies = false, hitv = false, test_pending = false;
    uint32_t slot = 0;
    v3 o = V3(0.f, 0.f, 0.f), dn = o;
    float tn0 = 0.f, tn1 = 0.f, tn2 = 0.f, tf0 = 0.f, tf1 = 0.f, tf2 = 0.f;     // slab distances of the renormalised ray, per axis, ordered
    float t_lo = 0.f, length = 1.f, t0 = 0.f, t1 = 0.f, t_i = 0.f, xi = 0.f;
    float s_hit = 0.f, s_nohit = 0.f;                                           // optical-depth scalars beyond / below which the test is decided
    float t = 0.f, tb = -INFINITY, s = 0.f;                                     // the step in progress: next sample, end, running density sum
    PathRng rng; rng.reset(0, a.k0, a.k1);
    for (;;) {
        const uint32_t n_pull = __popc(__ballot_sync(PV_FULL, ms == MS_PULL)), n_step = __popc(__ballot_sync(PV_FULL, ms == MS_STEP));
        const uint32_t n_samp = __popc(__ballot_sync(PV_FULL, ms == MS_SAMPLE)), n_ref = __popc(__ballot_sync(PV_FULL, ms == MS_REFILL));
        if ((n_pull | n_step | n_samp | n_ref) == 0) break;
        uint32_t pick = MS_SAMPLE, best = n_samp * WF_W_SAMPLE;
        if (n_step * WF_W_STEP > best) { pick = MS_STEP; best = n_step * WF_W_STEP; }
        if (n_pull * WF_W_PULL > best) { pick = MS_PULL; best = n_pull * WF_W_PULL; }
        if (n_ref * WF_W_REFILL > best) { pick = MS_REFILL; best = n_ref * WF_W_REFILL; }
        if (ms != pick) continue;
        if (pick == MS_SAMPLE) {
#pragma unroll 1
            for (int u = 0; u < WF_SAMPLES_PER_PASS; ++u) {
                const v3 P = med_to_volume_p(med, ray_at(o, dn, t));
                s += pow2 ? grid_density_pow2(med, P, ix, iy, iz) : grid_density(med, P);
                c_dens++;
                t += istep;
                if (!(t < tb)) { test_pending = true; ms = MS_STEP; break; }
            }
        } else if (pick == MS_STEP) {
            if (test_pending) {
                // the step has all its samples: xi > Tr.y() ?  (y(exp(-sig_t s)) lies between exp(-st_max s) y1 and exp(-st_min s) y1)
                test_pending = false;
                const float tau = homog ? s : s * istep;
                if (tau < s_nohit) hitv = false;
                else if (tau > s_hit) hitv = true;
                else {
                    const float elo = expf(-(st_max * tau)) * y_one, ehi = expf(-(st_min * tau)) * y_one;
                    if (xi > ehi * 1.0001f) hitv = true;
                    else if (xi < elo * 0.9999f) hitv = false;
                    else {
                        float yy = 0.f;
#pragma unroll 1
                        for (int b = 0; b < PV_NSPEC; ++b) yy += s_cie[b] * expf(-(s_st[b] * tau));
                        hitv = xi > __fdiv_rn(yy * 300.f, 106.856895f * (float)PV_NSPEC);
                    }
                }
                bool ended = hitv;
                if (!hitv) { t0 += a.stepsize; ended = !(t0 < t1); }
                if (ended) { ms = MS_PULL; continue; }                        // the segment's results are written when the lane pulls
            }
            if (rng.pos == 4) { ms = MS_REFILL; continue; }
            // ---- the next step: offset = RandomFloat() (Transmittance with sample == NULL), then BBox::IntersectP
            // (core/geometry.cpp:68-86) over [t_i * length, t0 * length] from the slab distances of the segment
            const float uo = rng.next();
            if (homog || length == 0.f) {
                uint32_t nsmp = 0;
                s = med_tau_scalar(med, o, dn, t_i, t0, istep, uo, &nsmp);
                test_pending = true;                                          // no samples to take: straight to the test
            } else {
                float ta = t_lo; tb = t0 * length;
                bool ok = true;
                ta = tn0 > ta ? tn0 : ta; tb = tf0 < tb ? tf0 : tb; ok = ok && !(ta > tb);
                ta = tn1 > ta ? tn1 : ta; tb = tf1 < tb ? tf1 : tb; ok = ok && !(ta > tb);
                ta = tn2 > ta ? tn2 : ta; tb = tf2 < tb ? tf2 : tb; ok = ok && !(ta > tb);
                s = 0.f;
                t = ta + uo * istep;
                if (ok && t < tb) ms = MS_SAMPLE; else test_pending = true;
            }
        } else if (pick == MS_REFILL) {
            const uint4 r = path_philox_block(rng.c0, rng.c1, rng.j++, rng.k0, rng.k1);
            rng.buf[0] = r.x; rng.buf[1] = r.y; rng.buf[2] = r.z; rng.buf[3] = r.w; rng.pos = 0;
            ms = MS_STEP;
        } else {
            // ---- PULL: the segment just ended (if any) hands its result over, the next one comes from the queue
            if (have) {
                if (!hitv && dies) w.state[slot] = WS_NEW;                    // no interaction before a matte surface, surface maps off: the path is over
                else {
                    w.march[slot] = make_float4(t0, t1, t_i, xi);
                    rng_store(w, slot, rng);
                    w.state[slot] = hitv ? WS_HIT : WS_MISS;
                    event_push(w, slot);
                }
                have = false;
            }
            cg::coalesced_group g = cg::coalesced_threads();
            unsigned int qi = 0;
            if (g.thread_rank() == 0) qi = atomicAdd(&w.ctr[1], (unsigned int)g.size());
            qi = g.shfl(qi, 0) + g.thread_rank();
            if (qi >= n_q) { ms = MS_DONE; continue; }
            slot = w.queue[qi];
            dies = (slot >> 31) != 0; slot &= 0x7fffffffu;
            const float4 f0 = w.frame[WF_AT(w.P, slot, 0)], f1 = w.frame[WF_AT(w.P, slot, 1)], m = w.march[slot];
            o = V3(f0.x, f0.y, f0.z);
            const v3 d = V3(f1.x, f1.y, f1.z);
            const v3 rnd = vdiv(d, vlen(d));
            t0 = m.x; t1 = m.y; t_i = m.z; xi = m.w;
            rng_load(w, slot, w.path[slot], a.k0, a.k1, rng);
            // the ray-only part of DensityRegion::tau for Ray(o, rnd, t_i, .): renormalisation and BBox::IntersectP's slabs
            length = vlen(rnd);
            dn = (homog || length == 0.f) ? rnd : vdiv(rnd, length);          // (the two cases that go through med_tau_scalar keep the ray as it is)
            const v3 po = med_to_volume_p(med, o), dv = med_to_volume_v(med, dn);
            float tn[3], tf[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float inv = __fdiv_rn(1.f, vcomp(dv, i));
                float tNear = (med.p0[i] - vcomp(po, i)) * inv, tFar = (med.p1[i] - vcomp(po, i)) * inv;
                if (tNear > tFar) { const float tt = tNear; tNear = tFar; tFar = tt; }
                tn[i] = tNear; tf[i] = tFar;
            }
            tn0 = tn[0]; tn1 = tn[1]; tn2 = tn[2]; tf0 = tf[0]; tf1 = tf[1]; tf2 = tf[2];
            t_lo = t_i * length;
            // tau above which xi > exp(-st_min tau) y1 * 1.0001 for certain, below which xi < exp(-st_max tau) y1 * 0.9999 for certain
            // (margins 1e-3 relative + 1e-4 absolute in the exponent, four orders above the rounding of expf): between them the test
            // is evaluated as written
            const float Lh = -logf(xi / (y_one * 1.0001f)), Ln = logf(y_one * 0.9999f / xi);
            s_hit = st_min > 0.f ? (Lh + 1e-3f * fabsf(Lh) + 1e-4f) / st_min : INFINITY;
            s_nohit = st_max > 0.f ? (Ln - 1e-3f * fabsf(Ln) - 1e-4f) / st_max : (Ln > 2e-4f ? INFINITY : -INFINITY);
            if (!(xi > 0.f) || !(s_hit == s_hit) || !(s_nohit == s_nohit)) { s_hit = INFINITY; s_nohit = -INFINITY; }   // degenerate xi: always as written
            have = true; test_pending = false; hitv = false;
            ms = MS_STEP;
        }
    }
    uint32_t cnt[3] = {0, 0, c_dens};
    flush_stats(a.stats, cnt, 3);
}

// ---------------------------------------------------------------------------------------------------------------- event
// Continuation stack of a slot: level `sp` of Frame storage.  Levels below WF_FAST_LEVELS are SoA over the slots; the levels above
// live in a page the slot takes from a pool the first time it gets that deep.  Returns false when the pool is empty: the frame is
// dropped, ctr[7] is raised and the host replays the (deterministic) wave with a larger pool.
__device__ __forceinline__ bool stack_store(const WaveState &w, uint32_t slot, int sp, const Frame &f) {
    if (sp < WF_FAST_LEVELS) { frame_store(w.stack + (size_t)sp * WF_FRAME_F4S(w.P), w.P, slot, f); return true; }
    uint32_t page = w.deep_page[slot];
    if (page == 0xffffffffu) {
        page = atomicAdd(&w.ctr[6], 1u);
        if (page >= w.deep_pages) { w.ctr[7] = 1u; return false; }
        w.deep_page[slot] = page;
    }
    frame_store(w.deep + ((size_t)page * (SH_MAXDEPTH - WF_FAST_LEVELS) + (size_t)(sp - WF_FAST_LEVELS)) * WF_STRIDE, 1u, 0u, f);
    return true;
}
__device__ __forceinline__ void stack_load(const WaveState &w, uint32_t slot, int sp, Frame &f) {
    if (sp < WF_FAST_LEVELS) { frame_load(w.stack + (size_t)sp * WF_FRAME_F4S(w.P), w.P, slot, f); return; }
    const uint32_t page = w.deep_page[slot];
    frame_load(w.deep + ((size_t)page * (SH_MAXDEPTH - WF_FAST_LEVELS) + (size_t)(sp - WF_FAST_LEVELS)) * WF_STRIDE, 1u, 0u, f);
}

template <class MV>
static __device__ __noinline__ float wf_tau(const MV m, float ox, float oy, float oz, float dx, float dy, float dz, float mint, float maxt,
                                            float stepSize, float u, uint32_t *nsamples) {
    return med_tau_scalar(m, V3(ox, oy, oz), V3(dx, dy, dz), mint, maxt, stepSize, u, nsamples);
}

template <bool SURF, int KIND>
__global__ void __launch_bounds__(WF_THREADS, WF_EVENT_MIN_CTAS) wf_event_kernel(ShootArgs a, WaveState w) {
    typedef typename std::conditional<(KIND & 2) != 0, MedView, MedViewPlain>::type MV;
    __shared__ float s_cie[PV_NSPEC], s_sa[PV_NSPEC], s_ss[PV_NSPEC], s_st[PV_NSPEC];
    const DevScene &sc = *a.sc;
    const DevMedium &gmed = sc.med;
    const MV med = make_medview<MV>(gmed);
    if (threadIdx.x < PV_NSPEC) {
        s_cie[threadIdx.x] = sc.cie_y[threadIdx.x]; s_sa[threadIdx.x] = gmed.sigma_a[threadIdx.x];
        s_ss[threadIdx.x] = gmed.sigma_s[threadIdx.x]; s_st[threadIdx.x] = gmed.sigma_a[threadIdx.x] + gmed.sigma_s[threadIdx.x];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { w.ctr[0] = 0; w.ctr[1] = 0; w.ctr[2] = 0; }      // the queue of the generation that starts here
    __syncthreads();
    // the slots the previous generation's trace and march kernels queued for this kernel
    const uint32_t half = (w.gen + 1u) & 1u, idx = blockIdx.x * WF_THREADS + threadIdx.x;
    const bool mine = idx < w.ctr[4 + half];
    const uint32_t slot = mine ? w.equeue[(size_t)half * w.P + idx] : 0u;
    const uint32_t st = mine ? w.state[slot] : (uint32_t)WS_IDLE;
    uint32_t cnt[5] = {0, 0, 0, 0, 0};            // -, -, density samples, -, overflows
    if (st == WS_HIT || st == WS_MISS || st == WS_POP) {
        Frame cur;
        frame_load(w.frame, w.P, slot, cur, &sc);
        const uint64_t path = w.path[slot];
        const uint64_t gblock = (path - 1) / SH_BLOCK + 1;
        PathRng rng;
        rng_load(w, slot, path, a.k0, a.k1, rng);
        const uint2 mi = w.misc[slot];
        uint32_t dep_seq = mi.x; int sp = (int)mi.y;
        int state = ST_SURFACE;
        bool pop = false;
        if (st == WS_POP) pop = true;
        else if (st == WS_HIT) {
            // ---- the medium interaction at t0 (photonshooter.cpp:82-128)
            const v3 o = V3(cur.o[0], cur.o[1], cur.o[2]), d = V3(cur.d[0], cur.d[1], cur.d[2]);
            const v3 rnd = vdiv(d, vlen(d));
            const float t0 = w.march[slot].x;
            const v3 pt = ray_at(o, rnd, t0);
            uint32_t ns = 0;
            const float dens = med_density(med, pt, &ns);
            cnt[2] += 2 * ns;
            float ys = 0.f, ya = 0.f;
#pragma unroll 1
            for (int b = 0; b < PV_NSPEC; ++b) { ys += s_cie[b] * (s_ss[b] * dens); ya += s_cie[b] * (s_sa[b] * dens); }
            ys = __fdiv_rn(ys * 300.f, 106.856895f * (float)PV_NSPEC); ya = __fdiv_rn(ya * 300.f, 106.856895f * (float)PV_NSPEC);
            const bool scatter = rng.next() > __fdiv_rn(ys, ya + ys);      // Q1 (photonshooter.cpp:88)
            if (!scatter) pop = true;
            else if (SURF && (a.flags & SF_VOLUME_DONE)) {
                // `if (scatter && !volumeDone)` (:96): with the volume map full the event is ignored and the surface code
                // runs with the unscattered ray
                cur.loop_i = -1;
            } else {
                if (cur.nI > 1) {
                    // ---- deposit (photonshooter.cpp:98-102), normalised by nshot of its block (:333)
                    deposit_photon(a, PC_VOLUME, gblock, path, dep_seq, pt, rnd, cur.alpha, (float)(gblock * SH_BLOCK));
                    dep_seq++;
                } else if (SURF) atomicAdd(&a.block_counts[PC_COUNT * a.wave_blocks + (uint32_t)(gblock - a.first_block)], 1u);   // shooter->nVolumePaths++ (:104)
                const float u1 = rng.next(), u2 = rng.next();
                const v3 dir = uniform_sample_sphere(u1, u2);
                const float pdf = __fdiv_rn(1.f, 4.f * PV_PI_F);
                const float ref = med_phase(med, pt, rnd, dir);
                if (ref == 0.f) pop = true;
                else {
#pragma unroll 1
                    for (int b = 0; b < PV_NSPEC; ++b) cur.alpha[b] = __fdiv_rn(cur.alpha[b] * ref, pdf);
                    cur.o[0] = pt.x; cur.o[1] = pt.y; cur.o[2] = pt.z; cur.d[0] = dir.x; cur.d[1] = dir.y; cur.d[2] = dir.z;
                    cur.mint = 0.f; cur.maxt = INFINITY; cur.loop_i = -1;
                    // Q2: after the scattered sub-path, the surface code runs with this ray and the hit above
                    if (sp < SH_MAXDEPTH && stack_store(w, slot, sp, cur)) sp++; else cnt[4]++;
                    state = ST_TRACE;                                      // the recursive call itself: same ray
                }
            }
        } else cur.loop_i = -1;
        uint32_t final_state = WS_NEW;
        for (;;) {
            if (pop) {
                if (sp == 0) { final_state = WS_NEW; break; }
                --sp;
                stack_load(w, slot, sp, cur);
                state = ST_SURFACE; pop = false;
            }
            if (state == ST_TRACE) { final_state = WS_TRACE; break; }
            // ---- surface part (photonshooter.cpp:131-227)
            const v3 o = V3(cur.o[0], cur.o[1], cur.o[2]), d = V3(cur.d[0], cur.d[1], cur.d[2]);
            if (cur.loop_i < 0) {
                const float uo = rng.next();
                // alpha *= Transmittance(photonRay) (:133-135).  With the surface maps off a matte hit ends the path without ever
                // reading the weight again (see wf_trace_kernel); the draw above is all that is left of it
                if (SURF || sc.mats[sc.prim_mat[cur.prim]].type != PV_MAT_MATTE) {
                    uint32_t ns = 0;
                    const float s = wf_tau<MV>(med, o.x, o.y, o.z, d.x, d.y, d.z, cur.mint, cur.maxt, a.istep4, uo, &ns);
                    cnt[2] += ns;
#pragma unroll 1
                    for (int b = 0; b < PV_NSPEC; ++b) cur.alpha[b] *= expf(-(s_st[b] * s));
                }
                cur.loop_i = 0;
                if (SURF) {
                    // ---- surface deposits (photonshooter.cpp:147-189).  hasNonSpecular == matte with a non-black Kd
                    // (materials/matte.cpp:55); glass has only specular components.
                    const pv_material &dm = sc.mats[sc.prim_mat[cur.prim]];
                    const bool nonspec = dm.type == PV_MAT_MATTE && (sc.mat_flags[sc.prim_mat[cur.prim]] & PV_MATF_KD) != 0;
                    if (nonspec) {
                        int cls = -1;
                        if ((cur.spec & 1) && cur.nI > 1) { if (a.flags & SF_WANT_CAUSTIC) cls = PC_CAUSTIC; }
                        else if (cur.nI == 1 && (a.flags & SF_WANT_INDIRECT) && (a.flags & SF_FINAL_GATHER)) cls = PC_DIRECT;
                        else if (cur.nI > 1 && (a.flags & SF_WANT_INDIRECT)) cls = PC_INDIRECT;
                        if (cls >= 0) {
                            const v3 hp = V3(cur.ip[0], cur.ip[1], cur.ip[2]);
                            deposit_photon(a, (uint32_t)cls, gblock, path, dep_seq, hp, -d, cur.alpha, 1.f);
                            dep_seq++;
                            // radiance-photon site (:178-188): p, Faceforward(n, -d), rho_r = Kd (Lambertian::rho), rho_t = 0;
                            // the two BSDF::rho calls draw 2 x 2 x StratifiedSample2D(6 x 6) = 288 floats
                            if ((a.flags & SF_FINAL_GATHER) && rng.next() < .125f) {
                                v3 rn = V3(cur.inn[0], cur.inn[1], cur.inn[2]);
                                if (vdot(rn, -d) < 0.f) rn = -rn;
                                deposit_photon(a, PC_RADIANCE, gblock, path, dep_seq, hp, rn, dm.kd, 1.f);
                                dep_seq++;
                                rng.skip(288);
                            }
                        }
                    }
                }
                if (cur.nI >= a.max_depth) { pop = true; continue; }
            }
            const pv_material &mat = sc.mats[sc.prim_mat[cur.prim]];
            const v3 wo = -d;
            const v3 nn = V3(cur.inn[0], cur.inn[1], cur.inn[2]);
            const v3 sn = vnorm(V3(cur.idpdu[0], cur.idpdu[1], cur.idpdu[2]));        // BSDF frame, reflection.cpp:619-627
            const v3 tn = vcross(nn, sn);
            if (mat.type == PV_MAT_MATTE) {
                // Lambertian bounce (reflection.cpp:323-330,534-598).  With the surface maps off the path always dies here (Q6:
                // indirectDone && !specularPath), but frames still on the stack keep drawing from this path's stream, so the number
                // of draws consumed must match the reference: 3 for BSDFSample, then the Russian roulette draw only if the sample
                // is valid.
                const float u0 = rng.next(), u1 = rng.next(); rng.next();
                const bool kd_black = !(sc.mat_flags[sc.prim_mat[cur.prim]] & PV_MATF_KD);
                bool bounced = false;
                if (!kd_black) {
                    const v3 wol = V3(vdot(wo, sn), vdot(wo, tn), vdot(wo, nn));
                    v3 wil;
                    concentric_sample_disk(u0, u1, &wil.x, &wil.y);
                    wil.z = __fsqrt_rn(fmaxf(0.f, 1.f - wil.x * wil.x - wil.y * wil.y));
                    if (wol.z < 0.f) wil.z *= -1.f;
                    const float pdf = (wol.z * wil.z > 0.f) ? fabsf(wil.z) * PV_INV_PI_F : 0.f;
                    if (pdf != 0.f) {
                        const v3 wiW = V3(sn.x * wil.x + tn.x * wil.y + nn.x * wil.z, sn.y * wil.x + tn.y * wil.y + nn.y * wil.z,
                                          sn.z * wil.x + tn.z * wil.y + nn.z * wil.z);
                        if (vdot(wiW, nn) * vdot(wo, nn) > 0.f) {
                            if (!SURF) rng.next();                              // continueProb draw
                            else {
                                // anew = alpha * f * |wi.n| / pdf, f = Kd / pi; Russian roulette on y(anew) / y(alpha) (:204-213)
                                const float adn = fabsf(vdot(wiW, nn));
                                float ynew = 0.f, yold = 0.f;
#pragma unroll 1
                                for (int b = 0; b < PV_NSPEC; ++b) {
                                    const float an = __fdiv_rn((cur.alpha[b] * (mat.kd[b] * PV_INV_PI_F)) * adn, pdf);
                                    ynew += s_cie[b] * an; yold += s_cie[b] * cur.alpha[b];
                                }
                                ynew = __fdiv_rn(ynew * 300.f, 106.856895f * (float)PV_NSPEC); yold = __fdiv_rn(yold * 300.f, 106.856895f * (float)PV_NSPEC);
                                const float continueProb = fminf(1.f, __fdiv_rn(ynew, yold));
                                // specularPath &= false, then `indirectDone && !specularPath` ends the path (:216-219)
                                if (!(rng.next() > continueProb) && (a.flags & SF_WANT_INDIRECT)) {
                                    int npos = 0;
#pragma unroll 1
                                    for (int b = 0; b < PV_NSPEC; ++b) {
                                        cur.alpha[b] = __fdiv_rn(__fdiv_rn((cur.alpha[b] * (mat.kd[b] * PV_INV_PI_F)) * adn, pdf), continueProb);
                                        npos += cur.alpha[b] > 0.f ? 1 : 0;
                                    }
                                    cur.spec = npos == 1 ? 2 : 0;        // specularPath = false; alpha re-made: lambda = extractLambda()
                                    cur.o[0] = cur.ip[0]; cur.o[1] = cur.ip[1]; cur.o[2] = cur.ip[2];
                                    cur.d[0] = wiW.x; cur.d[1] = wiW.y; cur.d[2] = wiW.z;
                                    cur.mint = cur.ieps; cur.maxt = INFINITY; cur.loop_i = -1;
                                    // the loop over `spectrums` has one entry here (no transmission, no split): a tail call
                                    bounced = true;
                                }
                            }
                        }
                    }
                }
                if (bounced) state = ST_TRACE; else pop = true;
            } else {
                // glass: SpecularReflection + dispersive SpecularTransmission (materials/glass.cpp:42-59)
                const uint8_t mf = sc.mat_flags[sc.prim_mat[cur.prim]];
                const bool hasR = (mf & PV_MATF_KR) != 0, hasT = (mf & PV_MATF_KT) != 0;
                const int matching = (hasR ? 1 : 0) + (hasT ? 1 : 0);
                int nz = 0;
#pragma unroll 1
                for (int b = 0; b < PV_NSPEC; ++b) if (cur.alpha[b] > 0.f) nz++;
                // hasTransmission && alpha.lambda < 0 && primitive->dispersive() (photonshooter.cpp:140-145).  lambda is path STATE
                // (see Frame::spec): a monochromatic child whose one bin underflowed to zero in a dense medium is not split again
                // here, it goes on as a black photon with lambda = -1 and ends at the next dispersive face.
                const bool do_split = hasT && !(cur.spec & 2) && mat.vn > 0.f;
                bool spawned = false;
                for (;;) {
                    // next spectrum of the split (splitSpectrum core/spectrum.h:253-265): bins with c != 0, in order
                    int bin = -1;
                    if (do_split) {
                        int seen = 0;
#pragma unroll 1
                        for (int b = 0; b < PV_NSPEC; ++b) if (cur.alpha[b] != 0.f) { if (seen == cur.loop_i) { bin = b; break; } seen++; }
                        if (bin < 0) break;
                    } else if (cur.loop_i > 0) break;
                    cur.loop_i++;
                    const float u0 = rng.next(), u1 = rng.next(), uc = rng.next();
                    (void)u0; (void)u1;
                    if (matching == 0) continue;
                    const int which = min((int)floorf(uc * matching), matching - 1);
                    const bool pickT = hasR ? (which == 1) : true;
                    const v3 wol = V3(vdot(wo, sn), vdot(wo, tn), vdot(wo, nn));
                    v3 wil; const float F = fresnel_dielectric(wol.z, mat.index);
                    float fpdf = 1.f;
                    if (!pickT) wil = V3(-wol.x, -wol.y, wol.z);
                    else {
                        const bool entering = wol.z > 0.f;
                        float ei = 1.f, et = mat.index;
                        int lam = -1;
                        if (do_split) lam = 400 + bin * 10;                    // extractLambda: integer step (700-400)/29 == 10
                        else if (nz == 1) {
#pragma unroll 1
                            for (int b = 0; b < PV_NSPEC; ++b) if (cur.alpha[b] > 0.f) lam = 400 + b * 10; }
                        if (lam > 0 && mat.vn > 0.f) {                         // Cauchy, reflection.cpp:155-161
                            const float lmu = __fdiv_rn((float)lam, 1000.f);
                            const float B = (float)((double)__fdiv_rn(et - 1.f, mat.vn) * 0.52345);
                            const float A = (float)((double)et - ((double)B / 0.34522792));
                            et = (float)((double)A + (double)B / ((double)lmu * (double)lmu));
                        }
                        if (!entering) { const float t = ei; ei = et; et = t; }
                        const float sini2 = fmaxf(0.f, 1.f - wol.z * wol.z);
                        const float eta = __fdiv_rn(ei, et);
                        const float sint2 = eta * eta * sini2;
                        if (sint2 >= 1.f) continue;                            // total internal reflection: pdf stays 0
                        float cost = __fsqrt_rn(fmaxf(0.f, 1.f - sint2));
                        if (entering) cost = -cost;
                        wil = V3(eta * -wol.x, eta * -wol.y, cost);
                    }
                    if (matching > 1) fpdf = __fdiv_rn(fpdf, (float)matching);
                    const v3 wiW = V3(sn.x * wil.x + tn.x * wil.y + nn.x * wil.z, sn.y * wil.x + tn.y * wil.y + nn.y * wil.z,
                                      sn.z * wil.x + tn.z * wil.y + nn.z * wil.z);
                    const float adn = fabsf(vdot(wiW, nn));
                    float anew[PV_NSPEC], ynew = 0.f, yold = 0.f; bool fblack = true;
#pragma unroll 1
                    for (int b = 0; b < PV_NSPEC; ++b) {
                        const float ab = do_split ? (b == bin ? cur.alpha[b] : 0.f) : cur.alpha[b];
                        const float fb = pickT ? __fdiv_rn((1.f - F) * mat.kt[b], fabsf(wil.z)) : __fdiv_rn(F * mat.kr[b], fabsf(wil.z));
                        fblack = fblack && fb == 0.f;
                        anew[b] = __fdiv_rn((ab * fb) * adn, fpdf);
                        ynew += s_cie[b] * anew[b]; yold += s_cie[b] * ab;
                    }
                    if (fblack) continue;
                    ynew = __fdiv_rn(ynew * 300.f, 106.856895f * (float)PV_NSPEC); yold = __fdiv_rn(yold * 300.f, 106.856895f * (float)PV_NSPEC);
                    const float continueProb = fminf(1.f, __fdiv_rn(ynew, yold));
                    if (rng.next() > continueProb) continue;
                    if (!(cur.spec & 1) && !(SURF && (a.flags & SF_WANT_INDIRECT))) continue;   // indirectDone && !specularPath
                    // spawn the child; this frame resumes at loop_i afterwards
                    if (sp < SH_MAXDEPTH && stack_store(w, slot, sp, cur)) sp++; else cnt[4]++;
                    int npos = 0;
#pragma unroll 1
                    for (int b = 0; b < PV_NSPEC; ++b) { cur.alpha[b] = __fdiv_rn(anew[b], continueProb); npos += cur.alpha[b] > 0.f ? 1 : 0; }
                    cur.spec = (cur.spec & 1) | (npos == 1 ? 2 : 0);           // alpha re-made: lambda = extractLambda()
                    cur.o[0] = cur.ip[0]; cur.o[1] = cur.ip[1]; cur.o[2] = cur.ip[2];
                    cur.d[0] = wiW.x; cur.d[1] = wiW.y; cur.d[2] = wiW.z;
                    cur.mint = cur.ieps; cur.maxt = INFINITY; cur.loop_i = -1;
                    spawned = true;
                    break;
                }
                if (spawned) state = ST_TRACE; else pop = true;
            }
        }
        if (final_state == WS_TRACE) {
            frame_store(w.frame, w.P, slot, cur);
            rng_store(w, slot, rng);
            w.misc[slot] = make_uint2(dep_seq, (uint32_t)sp);
        }
        w.state[slot] = final_state;
    }
    flush_stats(a.stats, cnt, 5);
}

// ---------------------------------------------------------------------------------------------------------------- host
struct WaveBuffers { void *base = nullptr; size_t bytes = 0; };

// Runs the paths [0, n_local_blocks * 4096) of `a` through the wavefront on ctx->stream.  Returns after the last generation has
// been launched and found empty (the caller's event pair brackets the whole run).
int pvi_wavefront_run(pv_ctx *ctx, const ShootArgs &a, bool surf, int kind, bool *replay) {
    *replay = false;
    const uint64_t total = (uint64_t)a.n_local_blocks * SH_BLOCK;
    // slots in flight: more slots = fewer, larger generations (measured on config 3: 1 M 97 ms, 2 M 88 ms, 4 M 83 ms per 2 M photons);
    // about 1.2 KB of state per slot, so the count follows the free device memory
    uint64_t slots = 1u << 19;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) slots = free_b > (48ull << 30) ? 1u << 22 : free_b > (12ull << 30) ? 1u << 21 : 1u << 19;
    }
    if (const char *e = getenv("PV_WF_SLOTS")) slots = std::max<uint64_t>(1024, std::min<uint64_t>(1u << 24, strtoull(e, nullptr, 10)));    // tuning knob
    const uint32_t P = (uint32_t)std::min<uint64_t>((slots + WF_THREADS - 1) / WF_THREADS * WF_THREADS, (total + WF_THREADS - 1) / WF_THREADS * WF_THREADS);
    // carve the slot arrays out of one allocation
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_frame = take(sizeof(float4) * WF_FRAME_F4S(P)), o_stack = take(sizeof(float4) * WF_FRAME_F4S(P) * WF_FAST_LEVELS),
                 o_state = take(4 * (size_t)P), o_path = take(8 * (size_t)P), o_rb = take(16 * (size_t)P), o_rj = take(4 * (size_t)P),
                 o_misc = take(8 * (size_t)P), o_march = take(16 * (size_t)P), o_queue = take(4 * (size_t)P), o_equeue = take(8 * (size_t)P), o_ctr = take(64), o_dpage = take(4 * (size_t)P);
    uint64_t floor_pages = P / 16;
    if (const char *e = getenv("PV_WF_DEEP_PAGES")) {       // test knob: start from a tiny pool so that the replay path runs
        floor_pages = 1;
        if (ctx->wf_deep_pages == 0) ctx->wf_deep_pages = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
    }
    if (ctx->wf_deep_pages == 0) ctx->wf_deep_pages = 4096;
    const uint32_t deep_pages = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(ctx->wf_deep_pages, floor_pages), P);
    const size_t o_deep = take(sizeof(float4) * WF_STRIDE * (size_t)(SH_MAXDEPTH - WF_FAST_LEVELS) * deep_pages);
    int rc = pv_ensure(ctx, &ctx->wf, &ctx->wf_bytes, off); if (rc) return rc;
    char *b = (char *)ctx->wf;
    WaveState w;
    w.P = P; w.volume_only = surf ? 0u : 1u; w.frame = (float4 *)(b + o_frame); w.stack = (float4 *)(b + o_stack); w.state = (uint32_t *)(b + o_state); w.path = (uint64_t *)(b + o_path);
    w.rng_buf = (uint4 *)(b + o_rb); w.rng_jp = (uint32_t *)(b + o_rj); w.misc = (uint2 *)(b + o_misc); w.march = (float4 *)(b + o_march);
    w.deep = (float4 *)(b + o_deep); w.deep_page = (uint32_t *)(b + o_dpage); w.deep_pages = deep_pages;
    w.queue = (uint32_t *)(b + o_queue); w.equeue = (uint32_t *)(b + o_equeue); w.ctr = (unsigned int *)(b + o_ctr); w.gen = 0;
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(w.state, 0, 4 * (size_t)P, ctx->stream));           // WS_NEW
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(w.ctr, 0, 64, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(w.deep_page, 0xFF, 4 * (size_t)P, ctx->stream));

    static void (*const k_trace[4])(ShootArgs, WaveState) = {wf_trace_kernel<0>, wf_trace_kernel<1>, wf_trace_kernel<2>, wf_trace_kernel<3>};
    static void (*const k_march[4])(ShootArgs, WaveState) = {wf_march_kernel<0>, wf_march_kernel<1>, wf_march_kernel<2>, wf_march_kernel<3>};
    static void (*const k_event[2][4])(ShootArgs, WaveState) = {
        {wf_event_kernel<false, 0>, wf_event_kernel<false, 1>, wf_event_kernel<false, 2>, wf_event_kernel<false, 3>},
        {wf_event_kernel<true, 0>, wf_event_kernel<true, 1>, wf_event_kernel<true, 2>, wf_event_kernel<true, 3>}};
    int per_sm = 0;
    PV_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_march[kind], WF_MARCH_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const unsigned grid = P / WF_THREADS, grid_march = (unsigned)std::min<uint64_t>((uint64_t)ctx->sm_count * per_sm, (P + WF_MARCH_THREADS - 1) / WF_MARCH_THREADS);
    uint64_t gens = total / P + 2;                       // most paths end in the generation they start in
    for (int round = 0; round < 100000; ++round) {
        for (uint64_t g = 0; g < gens; ++g, ++w.gen) {
            k_event[surf ? 1 : 0][kind]<<<grid, WF_THREADS, 0, ctx->stream>>>(a, w);
            k_trace[kind]<<<grid, WF_THREADS, 0, ctx->stream>>>(a, w);
            k_march[kind]<<<grid_march, WF_MARCH_THREADS, 0, ctx->stream>>>(a, w);
        }
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        unsigned int h_ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(h_ctr, w.ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        if (h_ctr[7]) {                                    // the deep-stack pool ran dry: frames were dropped, the wave must be traced again
            if (deep_pages >= P) { ctx->err = "pv_shoot: continuation stacks do not fit"; return PV_ENOMEM; }
            ctx->wf_deep_pages = std::min<uint64_t>((uint64_t)deep_pages * 4, P);
            *replay = true;
            return PV_OK;
        }
        if (h_ctr[2] == 0) return PV_OK;
        gens = std::max<uint64_t>(2, (uint64_t)((double)h_ctr[2] / P * 2.0) + 1);
    }
    ctx->err = "pv_shoot: the wavefront did not drain";
    return PV_ECUDA;
}
