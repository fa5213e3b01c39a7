// pv_shoot.cu -- photon shooting (K1-K3): PhotonShootingTask::Run / followPhoton
// (core/photonshooter.cpp:47-357) as one persistent-thread sm_100a kernel.  Two instantiations: the volume-photon
// branch alone (surface maps off: diffuse bounces end the path, nothing but volume photons is stored) and the
// pass with the surface maps on (caustic / indirect / direct deposits, radiance-photon sites, :147-189).
//
// Each thread is a small state machine {NEWPATH, TRACE, SURFACE}; all lanes of a warp meet at the head of the
// same loop every iteration (path regeneration happens in-loop), so divergence is confined to one iteration.
// The reference's recursion -- including its quirks Q1 (inverted scatter test), Q2 (after a scattered sub-path
// returns, control falls through into the surface code with the scattered ray and the ORIGINAL hit), Q3
// (transmittance re-marched from the segment start at every step, fresh offset), Q5, Q6 -- is unrolled onto an
// explicit per-thread stack of continuation frames that preserves the depth-first draw order, so a path consumes
// its Philox stream in exactly the order the CPU oracle does (tests compare the photon sets one to one).
// Light paths are dealt in the reference's blocks of 4096; a deposit of global block b is divided by
// nshot = 4096*b (Q4), independent of the number of ranks.  Deposits are appended with warp-aggregated atomics
// (one atomicAdd per coalesced group) into SoA planes; the 30-bin alpha goes out as one 128-byte line (8 x float4).
#include <algorithm>
#include <type_traits>
#include <vector>
#include "pv_ctx.h"

#include <cstdlib>
#include <chrono>
#include "pv_shoot.cuh"
#ifndef SH_TAU_MED
#define SH_TAU_MED const MV
#endif
// One copy of DensityRegion::tau for the shooter's two call sites (free-flight march, surface transmittance)
template <class MV>
__device__ __noinline__ float shoot_tau(SH_TAU_MED m, float ox, float oy, float oz, float dx, float dy, float dz, float mint, float maxt,
                                        float stepSize, float u, uint32_t *nsamples) {
    return med_tau_scalar(m, V3(ox, oy, oz), V3(dx, dy, dz), mint, maxt, stepSize, u, nsamples);
}
// EXPERIMENT (SH_COOP_TAU=1, off by default; see DESIGN.md "Shooter experiments"): warp-cooperative DensityRegion::tau
// (core/volume.cpp:296-310) for a density-grid medium.  The lanes of a full warp pool the samples of all their pending tau
// evaluations: each owner lane enumerates its sample parameters t_j exactly as the scalar loop does (SH_SLOTS at a time) into
// its row of `vals`; the rows are flattened with a prefix sum and the lanes take the samples round-robin (owner by binary
// search over the prefix table, the owner's ray by shuffle, result written over t_j); each owner adds its row up IN ORDER.
// Same sample points, same summation order: bit-identical to med_tau_scalar.  Every lane of the warp must call.
#ifndef SH_SLOTS
#define SH_SLOTS 16
#endif
#ifndef SH_COOP_TAU
#define SH_COOP_TAU 0
#endif
template <class Med>
__device__ __forceinline__ float coop_tau(const Med &m, uint32_t lane, bool want, v3 o, v3 d, float mint, float maxt,
                                          float stepSize, float u, float *vals, uint32_t *pref, uint32_t *nsamples) {
    float tcur = 0.f, tend = 0.f, sum = 0.f;
    v3 dn = V3(0.f, 0.f, 0.f);
    bool has = false;
    if (want) {
        const float length = vlen(d);
        if (length != 0.f) {
            dn = vdiv(d, length);
            float t0, t1;
            if (med_intersectp(m, o, dn, mint * length, maxt * length, &t0, &t1)) { has = true; tcur = t0 + u * stepSize; tend = t1; }
        }
    }
    for (;;) {
        uint32_t n = 0;
        if (has) while (n < SH_SLOTS && tcur < tend) { vals[lane * SH_SLOTS + n] = tcur; tcur += stepSize; ++n; }
        uint32_t inc = n;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t t = __shfl_up_sync(PV_FULL, inc, off); if (lane >= (uint32_t)off) inc += t; }
        const uint32_t total = __shfl_sync(PV_FULL, inc, 31);
        if (total == 0) break;
        pref[lane] = inc - n;
        if (lane == 0) pref[32] = total;
        __syncwarp();
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < total;
            uint32_t owner = 0;
            if (valid) {
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) if (pref[owner + sft] <= i) owner += sft;
            }
            const float ox = __shfl_sync(PV_FULL, o.x, owner), oy = __shfl_sync(PV_FULL, o.y, owner), oz = __shfl_sync(PV_FULL, o.z, owner);
            const float dx = __shfl_sync(PV_FULL, dn.x, owner), dy = __shfl_sync(PV_FULL, dn.y, owner), dz = __shfl_sync(PV_FULL, dn.z, owner);
            if (valid) {
                float *slot = vals + owner * SH_SLOTS + (i - pref[owner]);
                *slot = grid_density(m, med_to_volume_p(m, ray_at(V3(ox, oy, oz), V3(dx, dy, dz), *slot)));
            }
        }
        __syncwarp();
        for (uint32_t k = 0; k < n; ++k) sum += vals[lane * SH_SLOTS + k];
        if (nsamples) *nsamples += n;
        __syncwarp();
    }
    return sum * stepSize;
}

// The 30-bin spectrum loops stay ROLLED: unrolled (with an IEEE division or an expf per bin) they made the kernel 12 k
// instructions (197 KB) and the instruction cache missed on a third of the fetches; rolled, alpha[] is indexed dynamically
// and lives in local memory, which the free-flight march -- where the time goes -- never touches.
#ifndef SH_BIN_UNROLL
#define SH_BIN_UNROLL 1
#endif
#define SH_STR(x) #x
#define SH_PRAGMA_UNROLL(n) _Pragma(SH_STR(unroll n))
#define SH_UNROLL_BINS SH_PRAGMA_UNROLL(SH_BIN_UNROLL)
#ifndef SH_MIN_CTAS
#define SH_MIN_CTAS 8                // 32 warps/SM at 64 registers: latency-bound on density taps, the spills cost less than the occupancy gains (measured 4: 40.9 ms, 6: 36.1 ms, 8: 35.2 ms)
#endif
// KIND: bit 0 = the scene holds sphere primitives, bit 1 = the medium is exponential.  Plain scenes (triangles; homogeneous,
// rainbow or grid medium) run an instantiation that carries no trace of the other cases.
template <bool SURF, int KIND>
__global__ void __launch_bounds__(SH_THREADS, SH_MIN_CTAS) shoot_kernel(ShootArgs a) {
    constexpr bool SPH = (KIND & 1) != 0;
    typedef typename std::conditional<(KIND & 2) != 0, MedView, MedViewPlain>::type MV;
    __shared__ float s_cie[PV_NSPEC], s_sa[PV_NSPEC], s_ss[PV_NSPEC], s_st[PV_NSPEC];
    __shared__ uint32_t s_perm[41];
    __shared__ float s_minmax[3];
#if SH_COOP_TAU
    __shared__ float s_tau_vals[SH_THREADS / 32][32 * SH_SLOTS];
    __shared__ uint32_t s_tau_pref[SH_THREADS / 32][33];
#endif
    const DevScene &sc = *a.sc;
    const DevMedium &gmed = sc.med;
    const MV med = make_medview<MV>(gmed);             // extent / grid dimensions / grid pointer in registers
    if (threadIdx.x < PV_NSPEC) {
        s_cie[threadIdx.x] = sc.cie_y[threadIdx.x]; s_sa[threadIdx.x] = gmed.sigma_a[threadIdx.x];
        s_ss[threadIdx.x] = gmed.sigma_s[threadIdx.x]; s_st[threadIdx.x] = gmed.sigma_a[threadIdx.x] + gmed.sigma_s[threadIdx.x];
    }
    if (threadIdx.x < 41) s_perm[threadIdx.x] = a.perm[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        float mn = INFINITY, mx = 0.f, y1 = 0.f;
SH_UNROLL_BINS
        for (int b = 0; b < PV_NSPEC; ++b) { mn = fminf(mn, s_st[b]); mx = fmaxf(mx, s_st[b]); y1 += s_cie[b]; }
        s_minmax[0] = mn; s_minmax[1] = mx; s_minmax[2] = __fdiv_rn(y1 * 300.f, 106.856895f * (float)PV_NSPEC);
    }
    __syncthreads();
    const float st_min = s_minmax[0], st_max = s_minmax[1], y_one = s_minmax[2];
    const uint64_t total = (uint64_t)a.n_local_blocks * SH_BLOCK;
    const uint32_t halton_base[6] = {2, 3, 5, 7, 11, 13};

    Frame cur;
    Frame stack[SH_MAXDEPTH];
    int sp = 0, state = ST_NEWPATH;
    PathRng rng; rng.reset(0, a.k0, a.k1);
    uint64_t path = 0, gblock = 0; uint32_t lblock = 0, dep_seq = 0;
    uint32_t c_nodes = 0, c_tris = 0, c_dens = 0, c_seg = 0, c_ovf = 0, c_paths = 0;

    for (;;) {
        // Warp-wide meeting point of every iteration (all paths below end in `continue` or fall through to here).  The
        // votes are what makes the lanes RE-CONVERGE: without them independent thread scheduling lets each lane run its own
        // state sequence alone (measured: 2.2 active lanes per instruction).  Rounds alternate: while any lane is in a short
        // state (NEWPATH, SURFACE) only those lanes run and the lanes already at TRACE wait; once every live lane is at
        // TRACE they all trace together -- the segment march, where nearly all the time goes, runs with full warps.
        if (__all_sync(PV_FULL, state == ST_DONE)) break;
        const bool short_round = __any_sync(PV_FULL, state == ST_NEWPATH || state == ST_SURFACE);
#if SH_COOP_TAU
        const uint32_t part = __ballot_sync(PV_FULL, state == ST_TRACE);      // the lanes of a TRACE round (all of them, except at the very end)
#endif
        if (state == ST_DONE || (short_round && state == ST_TRACE)) continue;
        if (state == ST_NEWPATH) {
            // ---- fetch a light path (warp-aggregated counter) and emit it: photonshooter.cpp:248-275
            cg::coalesced_group g = cg::coalesced_threads();
            unsigned long long w = 0;
            if (g.thread_rank() == 0) w = atomicAdd(a.work, (unsigned long long)g.size());
            w = g.shfl(w, 0) + g.thread_rank();
            if (w >= total) { state = ST_DONE; continue; }
            lblock = (uint32_t)(w / SH_BLOCK);
            gblock = a.b_start + (uint64_t)lblock * a.world;
            path = (gblock - 1) * SH_BLOCK + (w % SH_BLOCK) + 1;
            rng.reset(path, a.k0, a.k1);
            dep_seq = 0; sp = 0; c_paths++;
            float u[6];
            {
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
            if (l.type == PV_LIGHT_POINT) {                                // lights/point.cpp:80-88
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
            float ad = fabsf(vdot(rd, rd));                                // AbsDot(Nl, photonRay.d) with Nl == ray.d
            float den = pdf * lightPdf;
            bool black = true;
            int npos = 0;
SH_UNROLL_BINS
            for (int b = 0; b < PV_NSPEC; ++b) {
                float Le = l.type == PV_LIGHT_SPOT ? l.intensity[b] * scale : l.intensity[b];
                cur.alpha[b] = __fdiv_rn(Le * ad, den);
                black = black && (cur.alpha[b] == 0.f);
                npos += cur.alpha[b] > 0.f ? 1 : 0;
            }
            if (pdf == 0.f || black) continue;                             // stays in ST_NEWPATH
            cur.o[0] = ro.x; cur.o[1] = ro.y; cur.o[2] = ro.z; cur.d[0] = rd.x; cur.d[1] = rd.y; cur.d[2] = rd.z;
            cur.mint = 0.f; cur.maxt = INFINITY; cur.nI = 0; cur.spec = 1 | (npos == 1 ? 2 : 0); cur.loop_i = -1; cur.prim = -1;
            state = ST_TRACE;
            continue;
        }
        const v3 o = V3(cur.o[0], cur.o[1], cur.o[2]), d = V3(cur.d[0], cur.d[1], cur.d[2]);
        bool pop = false;
        if (state == ST_TRACE) {
            // ---- followPhoton head: intersect, march the medium (photonshooter.cpp:54-128)
            c_seg++;
            float thit = cur.maxt;
            BvhCounters bc = {0, 0};
            int prim = bvh_traverse<false, SPH>(sc, o, d, cur.mint, &thit, &bc);
            c_nodes += bc.nodes; c_tris += bc.tris;
#if SH_COOP_TAU
            bool marching = false, interaction = false;
            v3 rnd = V3(0.f, 0.f, 0.f);
            float t0 = 1.0f, t1 = 0.0f, t_i = 0.f, xi = 0.f;
#endif
            if (prim < 0) pop = true;
            else {
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
                cur.prim = prim; cur.ip[0] = hp.x; cur.ip[1] = hp.y; cur.ip[2] = hp.z;
                cur.inn[0] = nn.x; cur.inn[1] = nn.y; cur.inn[2] = nn.z;
                cur.idpdu[0] = dpdu.x; cur.idpdu[1] = dpdu.y; cur.idpdu[2] = dpdu.z;
                cur.ieps = eps;
                cur.maxt = thit;                                            // GeometricPrimitive::Intersect: r.maxt = thit
                cur.nI++;
                float length = vlen(d);
                if (length == 0.f) pop = true;
                else {
#if SH_COOP_TAU
                    rnd = vdiv(d, length);
                    if (!med_intersectp(med, o, rnd, cur.mint * length, cur.maxt * length, &t0, &t1)) { t0 = 1.0f; t1 = 0.0f; }
                    t0 += rng.next() * a.stepsize;
                    t_i = t0;
                    xi = rng.next();
                    marching = true;
                }
            }
            {
                {
                    const bool coop = med.type == PV_MEDIUM_GRID && part == PV_FULL;
                    for (;;) {
                        const bool step = marching && !interaction && t0 < t1;
                        if (coop ? !__any_sync(PV_FULL, step) : !step) break;
                        float uo = 0.f;
                        if (step) uo = rng.next();
                        uint32_t ns = 0;
                        float s;
                        if (coop) s = coop_tau(med, threadIdx.x & 31, step, o, rnd, t_i, t0, a.istep4, uo, s_tau_vals[threadIdx.x >> 5],
                                               s_tau_pref[threadIdx.x >> 5], &ns);
                        else s = shoot_tau(med, o.x, o.y, o.z, rnd.x, rnd.y, rnd.z, t_i, t0, a.istep4, uo, &ns);
                        c_dens += ns;
                        if (!step) continue;
#else
                    v3 rnd = vdiv(d, length);
                    float t0, t1;
                    if (!med_intersectp(med, o, rnd, cur.mint * length, cur.maxt * length, &t0, &t1)) { t0 = 1.0f; t1 = 0.0f; }
                    t0 += rng.next() * a.stepsize;
                    const float t_i = t0;
                    const float xi = rng.next();
                    bool interaction = false;
                    while (t0 < t1) {
                        float uo = rng.next();                               // Transmittance(sample == NULL): offset = RandomFloat()
                        uint32_t ns = 0;
                        float s = shoot_tau(med, o.x, o.y, o.z, rnd.x, rnd.y, rnd.z, t_i, t0, a.istep4, uo, &ns);
                        c_dens += ns;
#endif
                        // xi > Tr.y() ?  y(exp(-sig_t s)) lies between exp(-st_max s) y1 and exp(-st_min s) y1
                        bool hitv;
                        float elo = expf(-(st_max * s)) * y_one, ehi = expf(-(st_min * s)) * y_one;
                        if (xi > ehi * 1.0001f) hitv = true;
                        else if (xi < elo * 0.9999f) hitv = false;
                        else {
                            float yy = 0.f;
SH_UNROLL_BINS
                            for (int b = 0; b < PV_NSPEC; ++b) yy += s_cie[b] * expf(-(s_st[b] * s));
                            hitv = xi > __fdiv_rn(yy * 300.f, 106.856895f * (float)PV_NSPEC);
                        }
#if SH_COOP_TAU
                        if (hitv) { interaction = true; continue; }
#else
                        if (hitv) { interaction = true; break; }
#endif
                        t0 += a.stepsize;
                    }
#if SH_COOP_TAU
                    if (!marching) { /* no segment to march: pop is set */ } else
#endif
                    if (interaction) {
                        v3 pt = ray_at(o, rnd, t0);
                        uint32_t ns = 0;
                        float dens = med_density(med, pt, &ns);
                        c_dens += 2 * ns;
                        float ys = 0.f, ya = 0.f;
SH_UNROLL_BINS
                        for (int b = 0; b < PV_NSPEC; ++b) { ys += s_cie[b] * (s_ss[b] * dens); ya += s_cie[b] * (s_sa[b] * dens); }
                        ys = __fdiv_rn(ys * 300.f, 106.856895f * (float)PV_NSPEC); ya = __fdiv_rn(ya * 300.f, 106.856895f * (float)PV_NSPEC);
                        bool scatter = rng.next() > __fdiv_rn(ys, ya + ys);     // Q1 (photonshooter.cpp:88)
                        if (!scatter) pop = true;
                        else if (SURF && (a.flags & SF_VOLUME_DONE)) {
                            // `if (scatter && !volumeDone)` (:96): with the volume map full the event is ignored and the
                            // surface code runs with the unscattered ray
                            state = ST_SURFACE; cur.loop_i = -1;
                            continue;
                        } else {
                            if (cur.nI > 1) {
                                // ---- deposit (photonshooter.cpp:98-102), normalised by nshot of its block (:333)
                                deposit_photon(a, PC_VOLUME, gblock, path, dep_seq, pt, rnd, cur.alpha, (float)(gblock * SH_BLOCK));
                                dep_seq++;
                            } else if (SURF) atomicAdd(&a.block_counts[PC_COUNT * a.wave_blocks + (uint32_t)(gblock - a.first_block)], 1u);   // shooter->nVolumePaths++ (:104)
                            float u1 = rng.next(), u2 = rng.next();
                            v3 dir = uniform_sample_sphere(u1, u2);
                            const float pdf = __fdiv_rn(1.f, 4.f * PV_PI_F);
                            float ref = med_phase(med, pt, rnd, dir);
                            if (ref == 0.f) pop = true;
                            else {
SH_UNROLL_BINS
                                for (int b = 0; b < PV_NSPEC; ++b) cur.alpha[b] = __fdiv_rn(cur.alpha[b] * ref, pdf);
                                cur.o[0] = pt.x; cur.o[1] = pt.y; cur.o[2] = pt.z; cur.d[0] = dir.x; cur.d[1] = dir.y; cur.d[2] = dir.z;
                                cur.mint = 0.f; cur.maxt = INFINITY; cur.loop_i = -1;
                                // Q2: after the scattered sub-path, the surface code runs with this ray and the hit above
                                if (sp < SH_MAXDEPTH) stack[sp++] = cur; else c_ovf++;
                                // the recursive call itself: same ray, state TRACE
                                state = ST_TRACE;
                                continue;
                            }
                        }
                    } else {
                        state = ST_SURFACE; cur.loop_i = -1;
                        continue;
                    }
                }
            }
        } else {
            // ---- surface part (photonshooter.cpp:131-227); caustic/indirect/direct maps are off on this path
            if (cur.loop_i < 0) {
                float uo = rng.next();
                uint32_t ns = 0;
                float s = shoot_tau(med, o.x, o.y, o.z, d.x, d.y, d.z, cur.mint, cur.maxt, a.istep4, uo, &ns);
                c_dens += ns;
SH_UNROLL_BINS
                for (int b = 0; b < PV_NSPEC; ++b) cur.alpha[b] *= expf(-(s_st[b] * s));
                cur.loop_i = 0;
                if (SURF) {
                    // ---- surface deposits (photonshooter.cpp:147-189).  hasNonSpecular == matte with a non-black Kd
                    // (materials/matte.cpp:55); glass has only specular components.
                    const pv_material &dm = sc.mats[sc.prim_mat[cur.prim]];
                    bool nonspec = false;
                    if (dm.type == PV_MAT_MATTE) {
SH_UNROLL_BINS
                        for (int b = 0; b < PV_NSPEC; ++b) nonspec = nonspec || dm.kd[b] != 0.f;
                    }
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
                if (cur.nI >= a.max_depth) pop = true;
            }
            if (!pop) {
                const pv_material &mat = sc.mats[sc.prim_mat[cur.prim]];
                const v3 wo = -d;
                const v3 nn = V3(cur.inn[0], cur.inn[1], cur.inn[2]);
                const v3 sn = vnorm(V3(cur.idpdu[0], cur.idpdu[1], cur.idpdu[2]));        // BSDF frame, reflection.cpp:619-627
                const v3 tn = vcross(nn, sn);
                if (mat.type == PV_MAT_MATTE) {
                    // Lambertian bounce (reflection.cpp:323-330,534-598).  With the surface maps off the path always dies
                    // here (Q6: indirectDone && !specularPath), but frames still on the stack keep drawing from this path's
                    // stream, so the number of draws consumed must match the reference: 3 for BSDFSample, then the Russian
                    // roulette draw only if the sample is valid.
                    float u0 = rng.next(), u1 = rng.next(); rng.next();
                    bool kd_black = true;
SH_UNROLL_BINS
                    for (int b = 0; b < PV_NSPEC; ++b) kd_black = kd_black && mat.kd[b] == 0.f;
                    if (!kd_black) {
                        v3 wol = V3(vdot(wo, sn), vdot(wo, tn), vdot(wo, nn));
                        v3 wil;
                        concentric_sample_disk(u0, u1, &wil.x, &wil.y);
                        wil.z = __fsqrt_rn(fmaxf(0.f, 1.f - wil.x * wil.x - wil.y * wil.y));
                        if (wol.z < 0.f) wil.z *= -1.f;
                        float pdf = (wol.z * wil.z > 0.f) ? fabsf(wil.z) * PV_INV_PI_F : 0.f;
                        if (pdf != 0.f) {
                            v3 wiW = V3(sn.x * wil.x + tn.x * wil.y + nn.x * wil.z, sn.y * wil.x + tn.y * wil.y + nn.y * wil.z,
                                        sn.z * wil.x + tn.z * wil.y + nn.z * wil.z);
                            if (vdot(wiW, nn) * vdot(wo, nn) > 0.f) {
                                if (!SURF) rng.next();                              // continueProb draw
                                else {
                                    // anew = alpha * f * |wi.n| / pdf, f = Kd / pi; Russian roulette on y(anew) / y(alpha) (:204-213)
                                    const float adn = fabsf(vdot(wiW, nn));
                                    float ynew = 0.f, yold = 0.f;
SH_UNROLL_BINS
                                    for (int b = 0; b < PV_NSPEC; ++b) {
                                        const float an = __fdiv_rn((cur.alpha[b] * (mat.kd[b] * PV_INV_PI_F)) * adn, pdf);
                                        ynew += s_cie[b] * an; yold += s_cie[b] * cur.alpha[b];
                                    }
                                    ynew = __fdiv_rn(ynew * 300.f, 106.856895f * (float)PV_NSPEC); yold = __fdiv_rn(yold * 300.f, 106.856895f * (float)PV_NSPEC);
                                    const float continueProb = fminf(1.f, __fdiv_rn(ynew, yold));
                                    // specularPath &= false, then `indirectDone && !specularPath` ends the path (:216-219)
                                    if (!(rng.next() > continueProb) && (a.flags & SF_WANT_INDIRECT)) {
                                        int npos = 0;
SH_UNROLL_BINS
                                        for (int b = 0; b < PV_NSPEC; ++b) {
                                            cur.alpha[b] = __fdiv_rn(__fdiv_rn((cur.alpha[b] * (mat.kd[b] * PV_INV_PI_F)) * adn, pdf), continueProb);
                                            npos += cur.alpha[b] > 0.f ? 1 : 0;
                                        }
                                        cur.spec = npos == 1 ? 2 : 0;        // specularPath = false; alpha re-made: lambda = extractLambda()
                                        cur.o[0] = cur.ip[0]; cur.o[1] = cur.ip[1]; cur.o[2] = cur.ip[2];
                                        cur.d[0] = wiW.x; cur.d[1] = wiW.y; cur.d[2] = wiW.z;
                                        cur.mint = cur.ieps; cur.maxt = INFINITY; cur.loop_i = -1;
                                        // the loop over `spectrums` has one entry here (no transmission, no split): a tail call
                                        state = ST_TRACE;
                                        continue;
                                    }
                                }
                            }
                        }
                    }
                    pop = true;
                } else {
                    // glass: SpecularReflection + dispersive SpecularTransmission (materials/glass.cpp:42-59)
                    bool hasR = false, hasT = false;
SH_UNROLL_BINS
                    for (int b = 0; b < PV_NSPEC; ++b) { hasR = hasR || mat.kr[b] != 0.f; hasT = hasT || mat.kt[b] != 0.f; }
                    const int matching = (hasR ? 1 : 0) + (hasT ? 1 : 0);
                    int nz = 0;
SH_UNROLL_BINS
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
SH_UNROLL_BINS
                            for (int b = 0; b < PV_NSPEC; ++b) if (cur.alpha[b] != 0.f) { if (seen == cur.loop_i) { bin = b; break; } seen++; }
                            if (bin < 0) break;
                        } else if (cur.loop_i > 0) break;
                        cur.loop_i++;
                        float u0 = rng.next(), u1 = rng.next(), uc = rng.next();
                        (void)u0; (void)u1;
                        if (matching == 0) continue;
                        int which = min((int)floorf(uc * matching), matching - 1);
                        bool pickT = hasR ? (which == 1) : true;
                        v3 wol = V3(vdot(wo, sn), vdot(wo, tn), vdot(wo, nn));
                        v3 wil; float F = fresnel_dielectric(wol.z, mat.index);
                        float fpdf = 1.f;
                        if (!pickT) wil = V3(-wol.x, -wol.y, wol.z);
                        else {
                            bool entering = wol.z > 0.f;
                            float ei = 1.f, et = mat.index;
                            int lam = -1;
                            if (do_split) lam = 400 + bin * 10;                    // extractLambda: integer step (700-400)/29 == 10
                            else if (nz == 1) {
SH_UNROLL_BINS
                                for (int b = 0; b < PV_NSPEC; ++b) if (cur.alpha[b] > 0.f) lam = 400 + b * 10; }
                            if (lam > 0 && mat.vn > 0.f) {                         // Cauchy, reflection.cpp:155-161
                                float lmu = __fdiv_rn((float)lam, 1000.f);
                                float B = (float)((double)__fdiv_rn(et - 1.f, mat.vn) * 0.52345);
                                float A = (float)((double)et - ((double)B / 0.34522792));
                                et = (float)((double)A + (double)B / ((double)lmu * (double)lmu));
                            }
                            if (!entering) { float t = ei; ei = et; et = t; }
                            float sini2 = fmaxf(0.f, 1.f - wol.z * wol.z);
                            float eta = __fdiv_rn(ei, et);
                            float sint2 = eta * eta * sini2;
                            if (sint2 >= 1.f) continue;                            // total internal reflection: pdf stays 0
                            float cost = __fsqrt_rn(fmaxf(0.f, 1.f - sint2));
                            if (entering) cost = -cost;
                            wil = V3(eta * -wol.x, eta * -wol.y, cost);
                        }
                        if (matching > 1) fpdf = __fdiv_rn(fpdf, (float)matching);
                        v3 wiW = V3(sn.x * wil.x + tn.x * wil.y + nn.x * wil.z, sn.y * wil.x + tn.y * wil.y + nn.y * wil.z,
                                    sn.z * wil.x + tn.z * wil.y + nn.z * wil.z);
                        float adn = fabsf(vdot(wiW, nn));
                        float anew[PV_NSPEC], ynew = 0.f, yold = 0.f; bool fblack = true;
SH_UNROLL_BINS
                        for (int b = 0; b < PV_NSPEC; ++b) {
                            float ab = do_split ? (b == bin ? cur.alpha[b] : 0.f) : cur.alpha[b];
                            float fb = pickT ? __fdiv_rn((1.f - F) * mat.kt[b], fabsf(wil.z)) : __fdiv_rn(F * mat.kr[b], fabsf(wil.z));
                            fblack = fblack && fb == 0.f;
                            anew[b] = __fdiv_rn((ab * fb) * adn, fpdf);
                            ynew += s_cie[b] * anew[b]; yold += s_cie[b] * ab;
                        }
                        if (fblack) continue;
                        ynew = __fdiv_rn(ynew * 300.f, 106.856895f * (float)PV_NSPEC); yold = __fdiv_rn(yold * 300.f, 106.856895f * (float)PV_NSPEC);
                        float continueProb = fminf(1.f, __fdiv_rn(ynew, yold));
                        if (rng.next() > continueProb) continue;
                        if (!(cur.spec & 1) && !(SURF && (a.flags & SF_WANT_INDIRECT))) continue;   // indirectDone && !specularPath
                        // spawn the child; this frame resumes at loop_i afterwards
                        if (sp < SH_MAXDEPTH) stack[sp++] = cur; else c_ovf++;
                        int npos = 0;
SH_UNROLL_BINS
                        for (int b = 0; b < PV_NSPEC; ++b) { cur.alpha[b] = __fdiv_rn(anew[b], continueProb); npos += cur.alpha[b] > 0.f ? 1 : 0; }
                        cur.spec = (cur.spec & 1) | (npos == 1 ? 2 : 0);           // alpha re-made: lambda = extractLambda()
                        cur.o[0] = cur.ip[0]; cur.o[1] = cur.ip[1]; cur.o[2] = cur.ip[2];
                        cur.d[0] = wiW.x; cur.d[1] = wiW.y; cur.d[2] = wiW.z;
                        cur.mint = cur.ieps; cur.maxt = INFINITY; cur.loop_i = -1;
                        spawned = true;
                        break;
                    }
                    if (spawned) { state = ST_TRACE; continue; }
                    pop = true;
                }
            }
        }
        if (pop) {
            if (sp == 0) state = ST_NEWPATH;
            else { cur = stack[--sp]; state = ST_SURFACE; }
        }
    }
    // ---- counters
    unsigned long long vals[6] = {c_nodes, c_tris, c_dens, c_seg, c_ovf, c_paths};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        unsigned long long v = vals[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(PV_FULL, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&a.stats[i], v);
    }
}

// PV_TIMING=1: wall time of the host-side phases of shooting on stderr (tuning aid)
namespace {
struct PhaseTimer {
    const char *what; bool on; std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(const char *w) : what(w), on(getenv("PV_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char *phase) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[pv timing] %s: %s %.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}

// ---------------------------------------------------------------- host: MT19937 only to reproduce task 0's Halton tables
static void halton_tables_task0(uint32_t perm[41]) {
    // RNG rng(31 * taskNum) with taskNum == 0, then PermutedHalton(6, rng): core/rng.cpp:43-107, montecarlo.cpp:380-397
    uint32_t mt[624]; int mti;
    mt[0] = 0u;
    for (mti = 1; mti < 624; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
    auto next = [&]() -> uint32_t {
        static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
        uint32_t y;
        if (mti >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u]; }
            for (; kk < 623; kk++) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u]; }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu); mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    };
    const uint32_t bases[6] = {2, 3, 5, 7, 11, 13};
    uint32_t *p = perm;
    for (int d = 0; d < 6; ++d) {
        uint32_t b = bases[d];
        for (uint32_t i = 0; i < b; ++i) p[i] = i;
        for (uint32_t i = 0; i < b; ++i) { uint32_t other = i + (next() % (b - i)); std::swap(p[i], p[other]); }
        p += b;
    }
}

// One wave of blocks.  surf_flags < 0: the volume-only kernel, counts[n_blocks].  Otherwise the all-maps kernel with the
// done flags SF_* held constant over the wave, counts[PC_COUNT + 1][n_blocks] (last row: first-hit scatter events, :104).
static int shoot_wave(pv_ctx *ctx, uint64_t first_block, uint32_t n_blocks, const pv_shoot_params *prm, int surf_flags, uint32_t *counts,
                      pv_shoot_stats *stats) {
    const bool surf = surf_flags >= 0;
    const uint32_t n_cls = surf ? PC_COUNT + 1 : 1;
    if (!ctx->has_scene || ctx->hscene.med.type == PV_MEDIUM_NONE || ctx->hscene.n_lights == 0) {
        ctx->err = "pv_shoot: scene needs a medium and at least one light"; return PV_ESTATE;
    }
    if (first_block < 1 || prm->world < 1 || prm->rank >= prm->world) { ctx->err = "pv_shoot: bad block / rank arguments"; return PV_EINVAL; }
    if (!(prm->stepsize > 0.f) || !(prm->integrator_stepsize > 0.f)) { ctx->err = "pv_shoot: step sizes must be > 0"; return PV_EINVAL; }
    if (first_block == 1 && !surf) { ctx->n_photons = 0; ctx->built = false; }
    if (n_blocks == 0) return PV_OK;
    // blocks of this rank inside the wave: b with (b - 1) % world == rank
    uint64_t b_start = first_block + ((prm->rank + prm->world - ((first_block - 1) % prm->world)) % prm->world);
    uint64_t b_end = first_block + n_blocks;              // exclusive
    uint32_t n_local = b_start < b_end ? (uint32_t)((b_end - b_start + prm->world - 1) / prm->world) : 0;
    memset(counts, 0, sizeof(uint32_t) * n_blocks * n_cls);
    if (n_local == 0) return PV_OK;

    int rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, sizeof(uint32_t) * (size_t)n_blocks * n_cls + 64); if (rc) return rc;
    uint32_t *d_counts = (uint32_t *)ctx->io2;
    unsigned long long *d_nout = ctx->d_counters + 1, *d_work = ctx->d_counters + 2, *d_stats = ctx->d_counters + 8;
    // first guess of the capacity on top of what is already stored: the deposits per path seen in this context's earlier waves
    // (with a margin), else 8% (volume only) / 150% (all maps); never more than half of the free device memory -- a wave that
    // overflows its buffer is replayed with the exact size anyway
    double &yield = ctx->shoot_yield[surf ? 1 : 0];
    const double per_path = yield > 0. ? yield * 1.25 : (surf ? 1.5 : 0.08);
    uint64_t want_cap = ctx->n_photons + (uint64_t)((double)n_local * SH_BLOCK * per_path) + 65536;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const uint64_t room = ctx->cap_photons + (uint64_t)(free_b / 2) / 160;         // 160 B per photon over the four planes
            if (want_cap > room && room > ctx->n_photons + 65536) want_cap = room;
        }
    }
    PhaseTimer tm("shoot_wave");
    for (int attempt = 0; attempt < 3; ++attempt) {
        rc = pvi_reserve_photons(ctx, want_cap); if (rc) return rc;
        tm.lap("reserve");
        ShootArgs a;
        a.sc = ctx->dscene; a.b_start = b_start; a.n_local_blocks = n_local; a.world = prm->world; a.first_block = first_block;
        a.stepsize = prm->stepsize; a.istep4 = 4.f * prm->integrator_stepsize; a.max_depth = prm->max_photon_depth;
        a.k0 = (uint32_t)prm->seed; a.k1 = (uint32_t)(prm->seed >> 32);
        halton_tables_task0(a.perm);
        a.pos = ctx->d_pos; a.wi = ctx->d_wi; a.alpha32 = ctx->d_alpha; a.ids = ctx->d_ids;
        a.n_out = d_nout; a.cap = ctx->cap_photons; a.block_counts = d_counts; a.work = d_work; a.stats = d_stats;
        a.wave_blocks = n_blocks; a.flags = surf ? (uint32_t)surf_flags : 0u;
        unsigned long long init_n = ctx->n_photons, zero = 0;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_nout, &init_n, sizeof(init_n), cudaMemcpyHostToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_work, &zero, sizeof(zero), cudaMemcpyHostToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemsetAsync(d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * n_blocks * n_cls, ctx->stream));
        int per_sm = 0;
        const int kind = (ctx->hscene.n_spheres != 0 ? 1 : 0) | (ctx->hscene.med.type == PV_MEDIUM_EXPONENTIAL ? 2 : 0);
        static void (*const kerns[2][4])(ShootArgs) = {
            {shoot_kernel<false, 0>, shoot_kernel<false, 1>, shoot_kernel<false, 2>, shoot_kernel<false, 3>},
            {shoot_kernel<true, 0>, shoot_kernel<true, 1>, shoot_kernel<true, 2>, shoot_kernel<true, 3>}};
        void (*kern)(ShootArgs) = kerns[surf ? 1 : 0][kind];
        {   // CUDA loads a kernel lazily at its first launch; do it here so that the load does not sit between the two timing events
            cudaFuncAttributes fa;
            PV_CUDA_CHECK(ctx, cudaFuncGetAttributes(&fa, kern));
        }
        PV_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SH_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        uint64_t total = (uint64_t)n_local * SH_BLOCK;
        int blocks = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * per_sm, (total + SH_THREADS - 1) / SH_THREADS);
        static const bool megakernel = getenv("PV_SHOOT_MEGAKERNEL") != nullptr;        // A/B knob: the persistent-thread kernel of round 1
        PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        bool replay = false;
        if (megakernel) kern<<<blocks, SH_THREADS, 0, ctx->stream>>>(a);
        else { rc = pvi_wavefront_run(ctx, a, surf, kind, &replay); if (rc) return rc; }
        PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        unsigned long long h_nout = 0, h_stats[8];
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(&h_nout, d_nout, sizeof(h_nout), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(counts, d_counts, sizeof(uint32_t) * n_blocks * n_cls, cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        tm.lap("kernels + readback");
        if (replay) { --attempt; continue; }               // (the pool is at most quadrupled log4(P) times)
        if (h_nout > ctx->cap_photons) { want_cap = h_nout + 65536; continue; }      // too small: grow and replay the (deterministic) wave
        float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        yield = std::max(yield, (double)(h_nout - ctx->n_photons) / ((double)n_local * SH_BLOCK));
        ctx->n_photons = h_nout;
        if (stats) {
            stats->nodes_visited += h_stats[0]; stats->tri_tests += h_stats[1]; stats->density_samples += h_stats[2];
            stats->segments += h_stats[3]; stats->stack_overflows += h_stats[4]; stats->paths_local += h_stats[5];
            stats->seconds += ms * 1e-3;
        }
        return PV_OK;
    }
    ctx->err = "pv_shoot: could not size the photon buffer";
    return PV_ENOMEM;
}
int pvi_shoot_blocks(pv_ctx *ctx, uint64_t first_block, uint32_t n_blocks, const pv_shoot_params *prm, uint32_t *counts, pv_shoot_stats *stats) {
    return shoot_wave(ctx, first_block, n_blocks, prm, -1, counts, stats);
}

__global__ void id_keys_kernel(const uint64_t *__restrict__ ids, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = ids[i]; vals[i] = (uint32_t)i; }
}
int pvi_reserve_set(pv_ctx *ctx, PhotonSet *s, uint64_t n) {
    if (n <= s->cap) return PV_OK;
    pvi_free_set(s);
    uint64_t cap = std::max<uint64_t>(n, 1024);
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->pos, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->wi, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->alpha, cap * 32 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->ids, cap * sizeof(uint64_t)));
    s->cap = cap;
    return PV_OK;
}
void pvi_free_set(PhotonSet *s) {
    if (s->pos) cudaFree(s->pos);
    if (s->wi) cudaFree(s->wi);
    if (s->alpha) cudaFree(s->alpha);
    if (s->ids) cudaFree(s->ids);
    *s = PhotonSet();
}
__global__ void permute_photons_kernel(const uint32_t *__restrict__ order, uint64_t n, const float *__restrict__ pos, const float *__restrict__ wi,
                                       const float *__restrict__ alpha, const uint64_t *__restrict__ ids, float *__restrict__ pos_o,
                                       float *__restrict__ wi_o, float *__restrict__ alpha_o, uint64_t *__restrict__ ids_o) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t j = t >> 3; uint32_t sub = (uint32_t)(t & 7);
    if (j >= n) return;
    uint32_t src = order[j];
    *(reinterpret_cast<float4 *>(alpha_o + j * 32) + sub) = *(reinterpret_cast<const float4 *>(alpha + (uint64_t)src * 32) + sub);
    if (sub < 3) { pos_o[3 * j + sub] = pos[3 * (uint64_t)src + sub]; wi_o[3 * j + sub] = wi[3 * (uint64_t)src + sub]; }
    if (sub == 3) ids_o[j] = ids[src];
}

// Drop photons of blocks > last_block and order the rest by id = (path index << 16 | deposit ordinal): the photon
// set and its order then depend only on (scene, seed, target), not on thread scheduling or the number of ranks.
// split: the ids carry a class in their top bits (pv_shoot_maps); class 0 stays the context's photon set, the others
// move to ctx->surf[class - 1].
static int shoot_finish(pv_ctx *ctx, uint64_t last_block, bool split, bool bound_always = false) {
    uint64_t n = ctx->n_photons;
    ctx->built = false;
    if (split) for (int c = 0; c < 4; ++c) ctx->surf[c].n = 0;
    if (n == 0) return PV_OK;
    if (n > 0xFFFFFFF0ull) { ctx->err = "too many photons"; return PV_EINVAL; }
    size_t need = n * (2 * sizeof(uint64_t) + 2 * sizeof(uint32_t)) + 256;
    int rc = pv_ensure(ctx, &ctx->scratch, &ctx->scratch_bytes, need); if (rc) return rc;
    uint64_t *keys = (uint64_t *)ctx->scratch, *keys_tmp = keys + n;
    uint32_t *vals = (uint32_t *)(keys_tmp + n), *vals_tmp = vals + n;
    PhaseTimer tm("shoot_finish");
    id_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_ids, n, keys, vals);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    // highest path index bounds the key width
    uint64_t max_path = last_block ? last_block * SH_BLOCK : ~0ull >> 16;
    int bits = 16; while (bits < 64 && (max_path >> (bits - 16)) != 0) ++bits;
    bits = std::min(64, bits + 1);
    // photons of later blocks may still be present: they sort to the end because their path index is larger
    uint64_t *skeys; uint32_t *svals;
    rc = pvi_sort_pairs_u64(ctx, keys, vals, keys_tmp, vals_tmp, n, 64, &skeys, &svals); if (rc) return rc;
    (void)bits;
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));      // the probes below are blocking copies on another stream
    tm.lap("sort by id");
    // count survivors: ids <= (last_block * 4096) << 16 | 0xffff
    auto upper_bound = [&](uint64_t limit, uint64_t *out) -> int {  // first position with key > limit (few D2H probes)
        uint64_t probe = 0, lo = 0, hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            PV_CUDA_CHECK(ctx, cudaMemcpy(&probe, skeys + mid, sizeof(uint64_t), cudaMemcpyDeviceToHost));
            if (probe <= limit) lo = mid + 1; else hi = mid;
        }
        *out = lo;
        return PV_OK;
    };
    const uint64_t path_limit = last_block ? (((last_block * SH_BLOCK) << 16) | 0xffffull) : ((1ull << 60) - 1);
    uint64_t keep = n;
    if (last_block || split || bound_always) { rc = upper_bound(path_limit, &keep); if (rc) return rc; }
    if (split) {
        for (uint32_t c = 1; c < PC_COUNT; ++c) {
            uint64_t lo = 0, hi = 0;
            rc = upper_bound(((uint64_t)c << 60) - 1, &lo); if (rc) return rc;
            rc = upper_bound(((uint64_t)c << 60) | path_limit, &hi); if (rc) return rc;
            PhotonSet &ps = ctx->surf[c - 1];
            rc = pvi_reserve_set(ctx, &ps, hi - lo); if (rc) return rc;
            ps.n = hi - lo;
            if (ps.n) {
                permute_photons_kernel<<<(unsigned)((ps.n * 8 + 255) / 256), 256, 0, ctx->stream>>>(svals + lo, ps.n, ctx->d_pos, ctx->d_wi, ctx->d_alpha,
                                                                                                  ctx->d_ids, ps.pos, ps.wi, ps.alpha, ps.ids);
                PV_CUDA_CHECK(ctx, cudaGetLastError());
            }
        }
    }
    tm.lap("bounds");
    float *np, *nw, *na; uint64_t *ni;
    uint64_t cap = std::max<uint64_t>(keep, 1024);
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&np, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&nw, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&na, cap * 32 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ni, cap * sizeof(uint64_t)));
    tm.lap("malloc");
    if (keep) {
        permute_photons_kernel<<<(unsigned)((keep * 8 + 255) / 256), 256, 0, ctx->stream>>>(svals, keep, ctx->d_pos, ctx->d_wi, ctx->d_alpha, ctx->d_ids,
                                                                                         np, nw, na, ni);
        PV_CUDA_CHECK(ctx, cudaGetLastError());
    }
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    tm.lap("permute");
    cudaFree(ctx->d_pos); cudaFree(ctx->d_wi); cudaFree(ctx->d_alpha); cudaFree(ctx->d_ids);
    ctx->d_pos = np; ctx->d_wi = nw; ctx->d_alpha = na; ctx->d_ids = ni; ctx->cap_photons = cap; ctx->n_photons = keep;
    tm.lap("free");
    return PV_OK;
}
int pvi_shoot_finish(pv_ctx *ctx, uint64_t last_block) { return shoot_finish(ctx, last_block, false); }
// order the context's photon set by id and drop the records whose id is not a volume-photon id (class bits set: the slack of
// pv_allgather_photons carries id = ~0)
int pvi_sort_photons_by_id(pv_ctx *ctx) { return shoot_finish(ctx, 0, false, true); }

// Single-rank driver == PhotonShootingTask::Run's outer loop (photonshooter.cpp:245-356): waves of blocks until the
// running photon count reaches the target at some block M; give-up rule of :285-299.
int pvi_shoot(pv_ctx *ctx, uint64_t n_wanted, const pv_shoot_params *prm_in, pv_shoot_stats *stats) {
    pv_shoot_params prm = *prm_in;
    pv_shoot_stats st; memset(&st, 0, sizeof(st));
    if (prm.world != 1 || prm.rank != 0) { ctx->err = "pv_shoot: use pv_shoot_blocks/pv_shoot_finish when world > 1"; return PV_EINVAL; }
    uint64_t max_paths = prm.max_paths ? prm.max_paths : ((uint64_t)1 << 40);
    uint64_t block = 0, total = 0, last = 0;
    uint32_t wave = 64;
    std::vector<uint32_t> counts;
    bool done = false; int rc = PV_OK;
    if (n_wanted == 0) { ctx->n_photons = 0; ctx->built = false; if (stats) *stats = st; return PV_OK; }
    while (!done) {
        counts.assign(wave, 0);
        rc = pvi_shoot_blocks(ctx, block + 1, wave, &prm, counts.data(), &st); if (rc) return rc;
        for (uint32_t i = 0; i < wave; ++i) {
            uint64_t nshot_before = block * SH_BLOCK;
            // "Unable to store enough photons.  Giving up." (photonshooter.cpp:285-299, unsuccessful() :37-39)
            if (nshot_before > 500000 && total < n_wanted && (total == 0 || total < SH_BLOCK / 1024)) {
                ctx->n_photons = 0; ctx->err = "Unable to store enough photons.  Giving up."; rc = PV_ENOPHOTONS; done = true; break;
            }
            block++; total += counts[i];
            if (total >= n_wanted || block * SH_BLOCK >= max_paths) { last = block; done = true; break; }
        }
        if (!done) {
            // size the next wave from the observed yield, aiming a little past the target
            double per_block = std::max(1e-3, (double)total / (double)block);
            double remaining = (double)(n_wanted - total) / per_block;
            wave = (uint32_t)std::min<double>(std::max<double>(remaining * 1.03 + 8, 64), 262144);
            uint64_t left = (max_paths / SH_BLOCK > block) ? (max_paths / SH_BLOCK - block) : 1;
            wave = (uint32_t)std::min<uint64_t>(wave, left);
        }
    }
    if (rc == PV_OK) rc = pvi_shoot_finish(ctx, last);
    st.paths = last * SH_BLOCK; st.blocks = last; st.photons_local = ctx->n_photons;
    if (stats) *stats = st;
    return rc;
}

// All photon maps in one pass == PhotonShootingTask::Run's outer loop for ONE task (photonshooter.cpp:245-356).  The
// reference's done flags change how later paths behave (no more scattering in the medium once the volume map is full,
// :96; diffuse bounces end the path once the indirect map is full, :218), and they change at block ends.  A wave of
// blocks is traced with the flags held constant; the host then replays the reference's per-block bookkeeping over the
// per-class counts, and if a flag flips at block M inside the wave, the wave is rolled back and traced again up to M
// (it is deterministic), so that every later block sees the new flags.  Waves aim just short of the next expected flip.
// Several ranks: every rank traces the blocks b with (b - 1) % world == rank of each wave, the per-class, per-block counts are
// summed over ranks through `allreduce`, and every rank then replays the SAME bookkeeping on the same numbers -- flags, flips,
// roll-backs and the last block come out identical everywhere without any other communication.
int pvi_shoot_maps(pv_ctx *ctx, const pv_maps_params *mp, const pv_shoot_params *prm_in, pv_allreduce_u32_fn allreduce, void *user,
                   pv_maps_stats *out) {
    pv_shoot_params prm = *prm_in;
    pv_maps_stats ms; memset(&ms, 0, sizeof(ms));
    if (prm.world < 1 || prm.rank >= prm.world) { ctx->err = "pv_shoot_maps: bad rank / world"; return PV_EINVAL; }
    if (prm.world > 1 && !allreduce) { ctx->err = "pv_shoot_maps: world > 1 needs pv_shoot_maps_ranks with an all-reduce callback"; return PV_EINVAL; }
    const uint64_t wanted[3] = {mp->n_volume_wanted, mp->n_caustic_wanted, mp->n_indirect_wanted};     // by class id
    bool done[3] = {wanted[0] == 0, wanted[1] == 0, wanted[2] == 0};
    ctx->n_photons = 0; ctx->built = false; ctx->rad_valid = false;
    for (int c = 0; c < 4; ++c) { ctx->surf[c].n = 0; ctx->map_paths[c] = 0; }
    if (done[0] && done[1] && done[2]) { if (out) *out = ms; return PV_OK; }
    const uint64_t max_paths = prm.max_paths ? prm.max_paths : ((uint64_t)1 << 40);
    uint64_t block = 0, tot[PC_COUNT] = {0, 0, 0, 0, 0}, first_hits = 0;
    uint64_t paths[4] = {0, 0, 0, 0};                       // caustic, indirect, direct, volume
    uint32_t wave = 64;
    std::vector<uint32_t> counts;
    bool finished = false, aborted = false;
    int rc = PV_OK;
    auto unsuccessful = [](uint64_t needed, uint64_t found) { return found < needed && (found == 0 || found < SH_BLOCK / 1024); };
    while (!finished) {
        const int flags = (done[1] ? 0 : SF_WANT_CAUSTIC) | (done[2] ? 0 : SF_WANT_INDIRECT) | (done[0] ? SF_VOLUME_DONE : 0) |
                          (mp->final_gather ? SF_FINAL_GATHER : 0);
        const uint64_t n_before = ctx->n_photons, first = block + 1;
        counts.assign((size_t)wave * (PC_COUNT + 1), 0);
        rc = shoot_wave(ctx, first, wave, &prm, flags, counts.data(), &ms.shoot); if (rc) return rc;
        if (prm.world > 1 && allreduce(counts.data(), counts.size(), user) != 0) { ctx->err = "pv_shoot_maps: the all-reduce callback failed"; return PV_EINVAL; }
        bool flip = false;
        uint32_t used = 0;
        for (uint32_t i = 0; i < wave; ++i) {
            // "Unable to store enough photons.  Giving up." over all three wanted counts (:285-299)
            if (block * SH_BLOCK > 500000 && (unsuccessful(wanted[1], tot[1]) || unsuccessful(wanted[2], tot[2]) || unsuccessful(wanted[0], tot[0]))) {
                aborted = true; finished = true; break;
            }
            block++; used++;
            if (!done[2]) {                                  // :303-318
                paths[1] += SH_BLOCK; tot[2] += counts[(size_t)PC_INDIRECT * wave + i];
                paths[2] += SH_BLOCK; tot[3] += counts[(size_t)PC_DIRECT * wave + i];
                if (tot[2] >= wanted[2]) { done[2] = true; flip = true; }
            }
            if (!done[1]) {                                  // :320-328
                paths[0] += SH_BLOCK; tot[1] += counts[(size_t)PC_CAUSTIC * wave + i];
                if (tot[1] >= wanted[1]) { done[1] = true; flip = true; }
            }
            if (!done[0]) {                                  // :330-341
                paths[3] += SH_BLOCK; tot[0] += counts[(size_t)PC_VOLUME * wave + i];
                if (tot[0] >= wanted[0]) { done[0] = true; flip = true; }
            }
            tot[4] += counts[(size_t)PC_RADIANCE * wave + i];
            first_hits += counts[(size_t)PC_COUNT * wave + i];
            if ((done[0] && done[1] && done[2]) || block * SH_BLOCK >= max_paths) { finished = true; break; }
            if (flip) break;
        }
        if (aborted) break;
        if (flip && !finished && used < wave) {
            // roll the wave back and trace it again up to the block of the flip, with the flags it started with
            ctx->n_photons = n_before;
            counts.assign((size_t)used * (PC_COUNT + 1), 0);
            rc = shoot_wave(ctx, first, used, &prm, flags, counts.data(), &ms.shoot); if (rc) return rc;
            ms.replayed_blocks += used;                      // (the counts of the replay are the ones already booked: no second all-reduce)
        }
        if (!finished) {
            // next wave: 97% of the way to the nearest expected flip, from the yields seen so far
            double nearest = 1e30;
            for (int c = 0; c < 3; ++c)
                if (!done[c]) {
                    const double per_block = (double)tot[c] / (double)block;
                    nearest = std::min(nearest, per_block > 0 ? (double)(wanted[c] - tot[c]) / per_block : 4.0 * (double)block);
                }
            wave = (uint32_t)std::min<double>(std::max<double>(nearest * 0.97, 16), 65536);
            const uint64_t left = max_paths / SH_BLOCK > block ? max_paths / SH_BLOCK - block : 1;
            wave = (uint32_t)std::min<uint64_t>(wave, left);
        }
    }
    if (aborted) {
        ctx->n_photons = 0;
        for (int c = 0; c < 4; ++c) ctx->surf[c].n = 0;
        ctx->err = "Unable to store enough photons.  Giving up.";
        rc = PV_ENOPHOTONS;
    } else rc = shoot_finish(ctx, block, true);
    ms.nshot = block * SH_BLOCK; ms.blocks = block;
    ms.n_caustic_paths = paths[0]; ms.n_indirect_paths = paths[1]; ms.n_direct_paths = paths[2]; ms.n_volume_paths = paths[3] + first_hits;
    ms.shoot.paths = ms.nshot; ms.shoot.blocks = block; ms.shoot.photons_local = ctx->n_photons;
    ms.n[0] = ctx->n_photons;
    for (int c = 0; c < 4; ++c) { ms.n[c + 1] = ctx->surf[c].n; ctx->map_paths[c] = c == 3 ? ms.n_volume_paths : paths[c]; }
    if (out) *out = ms;
    return rc;
}
