// pv_cellgather.cu -- the photon lookups of a whole slice of march steps, batched by photon-grid cell.
//
// LPhoton (integrators/photonvolume.cpp:65-108) = KdTree::Lookup (core/kdtree.h:150-183) + PhotonProcess
// (core/photonshooter.h:186-203) + the flux sum, for EVERY march step of the slice at once.  The lookups of a ray's steps
// do not depend on one another (only the Lv/Tr recurrence is sequential, pv_gather.cu gather_kernel<true>), so they are
// reordered: cg_keys_kernel gives each step the key of the photon-grid cell its sample point falls into, the radix sort of
// pv_build.cu orders the steps by that key, and cellgather_kernel walks the sorted list in batches of 32 neighbouring
// steps, ONE WARP PER BATCH:
//
//   stage   the photons {x, y, z, index} (and wi) of the cells within maxdist of the batch's bounding box go to shared
//           memory ONCE for all 32 queries, one TMA bulk copy per contiguous cell row (cp.async.bulk -> UBLKCP, completion
//           on an mbarrier), rows in ascending Morton order = ascending photon order;
//   scan    lane == QUERY: every lane walks the staged candidates (a broadcast 128-bit shared-memory load per candidate),
//           computes the reference's unfused (dx*dx + dy*dy) + dz*dz against its own query and appends the accepted
//           candidate to its own list in shared memory -- no cross-lane traffic, 32 distance tests per ~13 instructions;
//   weigh   lane == query: phase function per accepted photon (from the staged wi), photon index resolved;
//   sum     when a list could overflow, and at the end of the batch: 8 lanes x float4 per 128-byte alpha line (128-bit
//           loads), eight queries at a time, longest lists first; each (query, bin) sum runs
//           over the query's photons in ascending photon order in one fma chain, so a step's result does not depend on
//           which other steps share its batch (bit-identical under any sharding of the rays);
//   finish  radius = largest accepted distance, estimate as in LPhoton's tail, one 128-byte row of L_ii per step.
//
// Exactness: the staged block is a superset of every query's search sphere (cell coordinates are monotone in the
// coordinate, the box is widened by maxdist + the grid margin), and each candidate is tested with the reference's
// arithmetic, so a step finds exactly the photons with d2 < r2.  A step with MORE than nused of them needs the k-nearest
// selection (ties by photon index): it is handed to the warp-per-step kernel (gather_lii_kernel) through an overflow list.
//
// K-NEAREST MODE (template parameter KNN; nused <= CG_KMAX, maxdist larger than the cells -- the regime of the reference's shipped
// scenes: "nused 50" inside a generous maxdist).  Same sort, staging and scan, but every lane searches its OWN trial radius -- a
// little more than the radius that holds nused photons at the density of the 27 cells around the query, capped at maxdist and at
// what the 5 x 5 rows around a cell cover -- and appends what it accepts to an unsorted list of (d2, map position) of its own in
// shared memory.  When a list could overflow, and at the end of the batch, every lane cuts its list down to its nused smallest
// entries: the threshold is found by bisection on the bit pattern of d2 (a counting loop per pass, the same code for all lanes;
// ties at the threshold by original photon index like core/kdtree.h:150-183 + PhotonProcess), the survivors keep their order and
// later candidates are tested against the threshold.  A lane that ends with nused photons has exactly the k nearest (everything
// closer than its k-th lies inside the trial sphere it scanned completely); a lane with fewer whose trial radius was already
// maxdist has exactly the photons within maxdist; a lane with fewer and a smaller trial radius goes to the warp-per-step kernel
// through the overflow list.  Accepted candidates arrive in ascending map order whatever else is staged, and the cut keeps that
// order: the sums run over a step's photons in map order, like the fixed-radius mode's.
#include <algorithm>
#include <cstdlib>
#include "pv_gather.cuh"

#ifndef CG_WARPS
#define CG_WARPS 2
#endif
#define CG_THREADS (CG_WARPS * 32)
#ifndef CG_MIN_CTAS
#define CG_MIN_CTAS 5
#endif
#ifndef CG_STAGE
#define CG_STAGE 64                      // candidates per TMA round
#endif
#ifndef CG_CAP
#define CG_CAP 24                        // accepted candidates per query between two sum phases (multiple of 4)
#endif
#ifndef CG_U
#define CG_U 8                           // scan unroll
#endif
#ifndef CG_XSPAN
#define CG_XSPAN 2                       // a sub-batch spans at most this many coarse cells along x
#endif
#define CG_NB 64                         // histogram bins of MODE 2 (they share the list's storage: CG_NB <= CG_KLIST)
#define CG_KMAX 64                       // largest nused of the k-nearest mode
#ifndef CG_KLIST
#define CG_KLIST 96
#endif
//                      // capacity of a lane's candidate list in that mode (>= CG_KMAX + CG_U)

struct CgArgs {
    MapView m;
    const DevScene *sc;
    const pv_ray *rays;
    const StepRec *steps;
    const uint32_t *order;               // march steps of the slice sorted by cell key
    unsigned long long total;
    float maxdist;
    uint32_t nused;
    float *lii;
    uint32_t *overflow;
    unsigned long long *counters;
    pv_gather_stats *stats;
    float knn_rmax;                      // k-nearest mode: upper bound of a trial radius
    float knn_trial;                     // ... and the factor on the radius expected to hold nused photons
    const unsigned long long *total_dev; // when set: the number of entries of `order` is read from the device (a list made by an earlier launch)
    uint32_t cnt_batch, cnt_overflow;    // which of `counters` this launch uses for its batch counter and for the steps it hands over
};

// per-warp shared memory
#define CG_OFF_POS 0                                                       // staged candidates {x, y, z, photon index}
#define CG_OFF_WI (CG_OFF_POS + (CG_STAGE + CG_U) * 16)                   // their wi
#define CG_OFF_LIDX (CG_OFF_WI + CG_STAGE * 16)                           // accepted candidates -> photon indices, [slot][lane]
#define CG_OFF_LW (CG_OFF_LIDX + CG_CAP * 32 * 4)                         // their phase-function weights, [slot][lane]
#define CG_OFF_PART (CG_OFF_LW + CG_CAP * 32 * 4)                         // running flux sums, [query][8 x float4]
#define CG_OFF_Q (CG_OFF_PART + 32 * 8 * 16)                              // fcnt, perm, qcnt, qmx, qdens, qstep, run_rs, run_len
#define CG_OFF_MBAR (CG_OFF_Q + 8 * 32 * 4)
#define CG_WARP_BYTES (CG_OFF_MBAR + 16)
#define CG_OFF_HEAP CG_WARP_BYTES                                          // k-nearest mode only: d2 bits [CG_KLIST][32], map position [CG_KLIST][32]
#define CG_WARP_BYTES_KNN (CG_OFF_HEAP + 2 * CG_KLIST * 32 * 4)
static_assert(CG_WARP_BYTES % 16 == 0 && CG_OFF_PART % 16 == 0 && CG_OFF_Q % 16 == 0, "16-byte alignment");
static_assert(CG_CAP % 4 == 0 && CG_CAP >= CG_U && CG_STAGE % CG_U == 0, "list capacity / unroll");

__device__ __forceinline__ int cg_ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float cg_unordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ float cg_warp_min(float v) { return cg_unordered(__reduce_min_sync(PV_FULL, cg_ordered(v))); }
__device__ __forceinline__ float cg_warp_max(float v) { return cg_unordered(__reduce_max_sync(PV_FULL, cg_ordered(v))); }

// One thread per march step: sort key = cell of the sample point in the photon grid (clamped like the photons' own keys).
__global__ void __launch_bounds__(256) cg_keys_kernel(GridParams g, const pv_ray *__restrict__ rays, const StepRec *__restrict__ steps,
                                                      unsigned long long total, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const float4 ra = __ldg(reinterpret_cast<const float4 *>(steps + s)), rb = __ldg(reinterpret_cast<const float4 *>(steps + s) + 1);
    uint32_t key = g.table_size - 1;                          // dead steps (behind a Russian-roulette stop): anywhere, never looked up
    if (!(ra.z >= PV_RR_DEAD)) {
        const uint32_t ri = __float_as_uint(rb.w);
        const v3 ro = V3(__ldg(&rays[ri].o[0]), __ldg(&rays[ri].o[1]), __ldg(&rays[ri].o[2]));
        const v3 rd = V3(__ldg(&rays[ri].d[0]), __ldg(&rays[ri].d[1]), __ldg(&rays[ri].d[2]));
        const v3 q = ray_at(ro, rd, ra.x);
        key = pv_cell_key(g.xbits, pv_cell_coord(q.x, g.origin[0], g.inv_hx, g.dims[0]), pv_cell_coord(q.y, g.origin[1], g.inv_h, g.dims[1]),
                          pv_cell_coord(q.z, g.origin[2], g.inv_h, g.dims[2]));
    }
    keys[s] = key; vals[s] = (uint32_t)s;
}

// MODE 0: fixed radius (lists + streaming sums).  MODE 1: k-nearest, per-lane candidate lists cut by bisection (nused <= CG_KMAX).
// MODE 2: k-nearest by RADIUS -- two passes over the staged block: the first counts every lane's candidates into a histogram of
// d2 (CG_NB bins over its trial sphere), from which the lane reads the bin b* its nused-th photon falls into; the second is the
// fixed-radius accumulation of everything in the bins below b*, while the few candidates IN bin b* go to a small list that is cut
// to the (nused - taken) nearest and added last.  Any nused; also what finishes the steps the fixed-radius mode hands over (more
// than nused photons within maxdist: the dense end of a shot map).
template <int MODE>
__global__ void __launch_bounds__(CG_THREADS, CG_MIN_CTAS) cellgather_kernel(CgArgs a) {
    constexpr bool KNN = MODE == 1, HIST = MODE == 2;
    extern __shared__ __align__(128) unsigned char cg_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = cg_smem + (size_t)warp * (MODE ? CG_WARP_BYTES_KNN : CG_WARP_BYTES);
    uint32_t *hd2 = reinterpret_cast<uint32_t *>(base + CG_OFF_HEAP) + lane;             // this lane's candidate list: entry e at [e * 32]
    uint32_t *hpos = reinterpret_cast<uint32_t *>(base + CG_OFF_HEAP + CG_KLIST * 32 * 4) + lane;
    float4 *spos = reinterpret_cast<float4 *>(base + CG_OFF_POS);
    const float4 *swi = reinterpret_cast<const float4 *>(base + CG_OFF_WI);
    uint32_t *lidx = reinterpret_cast<uint32_t *>(base + CG_OFF_LIDX);
    float *lw = reinterpret_cast<float *>(base + CG_OFF_LW);
    float4 *part = reinterpret_cast<float4 *>(base + CG_OFF_PART);
    uint32_t *fcnt = reinterpret_cast<uint32_t *>(base + CG_OFF_Q), *perm = fcnt + 32, *qcnt = perm + 32;
    float *qmx = reinterpret_cast<float *>(qcnt + 32), *qdens = qmx + 32;
    uint32_t *qstep = reinterpret_cast<uint32_t *>(qdens + 32), *run_rs = qstep + 32, *run_len = run_rs + 32;
    const uint32_t lidx_addr = smem_u32(lidx) + lane * 4u;
    const uint32_t mbar = smem_u32(base + CG_OFF_MBAR), spos_addr = smem_u32(spos), swi_addr = smem_u32(swi);
    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();
    uint32_t phase = 0;

    const GridParams &g = a.m.g;
    const DevMedium &med = a.sc->med;
    const uint32_t grp = lane >> 3, sub = lane & 7, gmask8 = 0xFFu << (grp * 8);
    float sg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { const uint32_t b = sub * 4 + c; sg[c] = b < PV_NSPEC ? med.sigma_s[b] : 0.f; }
    const float hg = med.g, pc = (1.f / (4.f * PV_PI_F)) * (1.f - hg * hg), gg1 = 1.f + hg * hg, g2 = 2.f * hg;
    const bool iso = hg == 0.f;
    const float r2 = a.maxdist * a.maxdist, slack_fixed = a.maxdist + g.margin;
    const float4 *a4 = reinterpret_cast<const float4 *>(a.m.alpha32) + sub;
    const uint32_t xspan = (uint32_t)CG_XSPAN << g.xshift;
    unsigned long long st_cand = 0;
    uint32_t st_lookups = 0, st_found = 0, st_heap = 0;
    const unsigned long long total = a.total_dev ? *a.total_dev : a.total;
    const unsigned long long nbatch = (total + 31ull) >> 5;

    for (;;) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(a.counters + a.cnt_batch, 1ull);
        b = __shfl_sync(PV_FULL, b, 0);
        if (b >= nbatch) break;
        const unsigned long long qi = b * 32ull + lane;
        bool valid = qi < total;
        uint32_t s = 0; float dens = 0.f;
        v3 q = V3(0.f, 0.f, 0.f), w = V3(0.f, 0.f, 0.f);
        if (valid) {
            s = __ldg(a.order + qi);
            const float4 ra = __ldg(reinterpret_cast<const float4 *>(a.steps + s)), rb = __ldg(reinterpret_cast<const float4 *>(a.steps + s) + 1);
            if (ra.z >= PV_RR_DEAD) valid = false;
            else {
                const uint32_t ri = __float_as_uint(rb.w);
                const v3 ro = V3(__ldg(&a.rays[ri].o[0]), __ldg(&a.rays[ri].o[1]), __ldg(&a.rays[ri].o[2]));
                const v3 rd = V3(__ldg(&a.rays[ri].d[0]), __ldg(&a.rays[ri].d[1]), __ldg(&a.rays[ri].d[2]));
                q = ray_at(ro, rd, ra.x); w = -rd; dens = ra.w;
            }
        }
        int cxf = 0; uint32_t row = 0;
        if (valid) {
            cxf = pv_cell_coord(q.x, g.origin[0], g.inv_hx, g.dims[0]);
            row = pv_morton2((uint32_t)pv_cell_coord(q.y, g.origin[1], g.inv_h, g.dims[1]), (uint32_t)pv_cell_coord(q.z, g.origin[2], g.inv_h, g.dims[2]));
        }
        // ---- k-nearest mode: this query's trial radius from the photon count of the 27 (coarse) cells around it
        float rq = a.maxdist, r2q = r2;
        uint32_t hc = 0; float hbound = INFINITY;                                  // list length; d2 of the K-th nearest so far once a cut has been made
        const uint32_t K = a.nused;
        uint32_t kcut = K;                                                         // how many entries of its list a lane keeps at a cut
        uint32_t bstar = CG_NB; float hscale = 0.f;                                // MODE 2: the lane's boundary bin, bins per unit of d2
        if ((KNN || (HIST && a.maxdist > g.h)) && valid) {
            const int cy = pv_cell_coord(q.y, g.origin[1], g.inv_h, g.dims[1]), cz = pv_cell_coord(q.z, g.origin[2], g.inv_h, g.dims[2]);
            const int cxc = cxf >> g.xshift;
            const int xlo = max((cxc - 1) << g.xshift, 0), xhi = min(((cxc + 2) << g.xshift) - 1, g.dims[0] - 1);
            uint32_t n27 = 0;
#pragma unroll 1
            for (int j = 0; j < 9; ++j) {
                const int y = cy + j % 3 - 1, z = cz + j / 3 - 1;
                if (y < 0 || y >= g.dims[1] || z < 0 || z >= g.dims[2]) continue;
                const uint32_t rowkey = pv_morton2((uint32_t)y, (uint32_t)z) << g.xbits;
                n27 += __ldg(a.m.cell_start + (rowkey | (uint32_t)xhi) + 1) - __ldg(a.m.cell_start + (rowkey | (uint32_t)xlo));
            }
            // radius holding K photons at that density: h * cbrt(27 * 3 K / (4 pi n27)); 20 % on top for its fluctuation
            const float rk = g.h * cbrtf(6.4458f * (float)K / (float)max(n27, 1u));
            rq = fminf(fminf(a.knn_trial * rk, a.knn_rmax), a.maxdist);
            r2q = rq < a.maxdist ? rq * rq : r2;
        }
        if (HIST) {
            hscale = r2q > 0.f ? (float)CG_NB / r2q : 0.f;
#pragma unroll 1
            for (int bq = 0; bq < CG_NB; ++bq) hd2[bq * 32] = 0u;                  // this lane's histogram (the storage becomes its boundary list in pass 2)
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        qstep[lane] = s; qdens[lane] = dens;
        uint32_t tot = 0, pad_idx = 0; float mx = 0.f; bool handed_over = false;
        uint32_t remaining = __ballot_sync(PV_FULL, valid);
        // Per-lane list of accepted candidates, [slot][lane]: entries [0, res) are resolved to (photon index, weight), entries
        // [res, cnt) still hold the candidate's position in the current stage.  The lists live on across stages and sub-batches
        // and are summed when one could overflow, and at the end of the batch.
        uint32_t lp = lidx_addr, res = 0;                                          // shared-memory address of this lane's next free slot
        // ---- weigh (lane == query): phase function of each newly accepted photon, photon index resolved
        auto resolve = [&]() {
            const uint32_t cnt = (lp - lidx_addr) >> 7;
            const uint32_t lo = __reduce_min_sync(PV_FULL, res), hi = __reduce_max_sync(PV_FULL, cnt);
            for (uint32_t e = lo; e < hi; ++e) {
                if (e >= res && e < cnt) {
                    const uint32_t c = lidx[e * 32 + lane];
                    float ph = pc;
                    if (!iso) {
                        const float4 wv = swi[c];
                        const float costheta = -(wv.x * w.x + wv.y * w.y + wv.z * w.z);      // Dot(wi, -w)
                        float rsq;                                                 // rsqrtf() of a normal number (the argument is >= (1 - |g|)^2)
                        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rsq) : "f"(gg1 - g2 * costheta));
                        ph = pc * rsq * rsq * rsq;
                    }
                    pad_idx = __float_as_uint(spos[c].w);
                    lidx[e * 32 + lane] = pad_idx;
                    lw[e * 32 + lane] = ph;
                }
            }
            res = cnt;
        };
        // ---- sum: 8 lanes x float4 per alpha line; each 8-lane group runs TWO queries' fma chains side by side
        auto sum = [&]() {
            const uint32_t cnt = (lp - lidx_addr) >> 7;
            const uint32_t longest = __reduce_max_sync(PV_FULL, cnt);
            // every list is padded to the same multiple of four with zero-weight entries of the query's own last photon (an
            // exact no-op in the sums)
            const uint32_t m4 = (longest + 3u) & ~3u;
            for (uint32_t e = __reduce_min_sync(PV_FULL, cnt); e < m4; ++e)
                if (e >= cnt) { lidx[e * 32 + lane] = pad_idx; lw[e * 32 + lane] = 0.f; }
            // queries in descending order of list length (in units of four entries, the granularity of the sums): the eight lists
            // summed side by side then have nearly equal lengths, so few lanes idle on padding.  Counting sort with one ballot
            // per possible length.
            const uint32_t c4n = (cnt + 3u) >> 2;
            uint32_t rank = 0, before = 0;
#pragma unroll
            for (int v = CG_CAP / 4; v >= 0; --v) {
                const uint32_t mv = __ballot_sync(PV_FULL, c4n == (uint32_t)v);
                if (c4n == (uint32_t)v) rank = before + __popc(mv & lanemask_lt());
                before += __popc(mv);
            }
            fcnt[lane] = cnt; perm[rank] = lane;
            __syncwarp();
            // the trip count of a round is warp-uniform: the longest of its eight lists is the first one
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                const uint32_t nq = (fcnt[perm[8 * i]] + 3u) & ~3u;
                if (nq == 0) break;
                const uint32_t qa = perm[8 * i + grp], qb = perm[8 * i + 4 + grp];
                float4 A = part[qa * 8 + sub], B = part[qb * 8 + sub];
                const uint32_t *ia = lidx + qa, *ib = lidx + qb; const float *wa = lw + qa, *wb = lw + qb;
#pragma unroll 1
                for (uint32_t e = 0; e < nq; e += 4, ia += 128, ib += 128, wa += 128, wb += 128) {
                    const uint32_t a0 = ia[0], a1 = ia[32], a2 = ia[64], a3 = ia[96], b0 = ib[0], b1 = ib[32], b2 = ib[64], b3 = ib[96];
                    const float4 x0 = __ldg(a4 + (size_t)a0 * 8), y0 = __ldg(a4 + (size_t)b0 * 8), x1 = __ldg(a4 + (size_t)a1 * 8),
                                 y1 = __ldg(a4 + (size_t)b1 * 8), x2 = __ldg(a4 + (size_t)a2 * 8), y2 = __ldg(a4 + (size_t)b2 * 8),
                                 x3 = __ldg(a4 + (size_t)a3 * 8), y3 = __ldg(a4 + (size_t)b3 * 8);
                    const float u0 = wa[0], u1 = wa[32], u2 = wa[64], u3 = wa[96], v0 = wb[0], v1 = wb[32], v2 = wb[64], v3 = wb[96];
                    A.x = fmaf(x0.x, u0, A.x); A.y = fmaf(x0.y, u0, A.y); A.z = fmaf(x0.z, u0, A.z); A.w = fmaf(x0.w, u0, A.w);
                    B.x = fmaf(y0.x, v0, B.x); B.y = fmaf(y0.y, v0, B.y); B.z = fmaf(y0.z, v0, B.z); B.w = fmaf(y0.w, v0, B.w);
                    A.x = fmaf(x1.x, u1, A.x); A.y = fmaf(x1.y, u1, A.y); A.z = fmaf(x1.z, u1, A.z); A.w = fmaf(x1.w, u1, A.w);
                    B.x = fmaf(y1.x, v1, B.x); B.y = fmaf(y1.y, v1, B.y); B.z = fmaf(y1.z, v1, B.z); B.w = fmaf(y1.w, v1, B.w);
                    A.x = fmaf(x2.x, u2, A.x); A.y = fmaf(x2.y, u2, A.y); A.z = fmaf(x2.z, u2, A.z); A.w = fmaf(x2.w, u2, A.w);
                    B.x = fmaf(y2.x, v2, B.x); B.y = fmaf(y2.y, v2, B.y); B.z = fmaf(y2.z, v2, B.z); B.w = fmaf(y2.w, v2, B.w);
                    A.x = fmaf(x3.x, u3, A.x); A.y = fmaf(x3.y, u3, A.y); A.z = fmaf(x3.z, u3, A.z); A.w = fmaf(x3.w, u3, A.w);
                    B.x = fmaf(y3.x, v3, B.x); B.y = fmaf(y3.y, v3, B.y); B.z = fmaf(y3.z, v3, B.z); B.w = fmaf(y3.w, v3, B.w);
                }
                part[qa * 8 + sub] = A; part[qb * 8 + sub] = B;
            }
            __syncwarp();
            tot += cnt; lp = lidx_addr; res = 0;
        };

        // ---- k-nearest mode: every lane cuts its list down to its K smallest (d2, original index) entries, order kept
        auto knn_cut = [&]() {
            const uint32_t nmax = __reduce_max_sync(PV_FULL, hc);
            const bool need = hc > kcut;
            if (!__any_sync(PV_FULL, need)) return;
            uint32_t lo = 0xffffffffu, hi = 0u;
            for (uint32_t e = 0; e < nmax; ++e) if (need && e < hc) { const uint32_t key = hd2[e * 32]; lo = min(lo, key); hi = max(hi, key); }
            // smallest threshold with at least K keys <= it (positive floats order like their bit patterns); a pass that counts
            // exactly K ends the search at once
            bool open = need && lo < hi;
            while (__any_sync(PV_FULL, open)) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                uint32_t c = 0;
                for (uint32_t e = 0; e < nmax; ++e) if (open && e < hc) c += hd2[e * 32] <= mid ? 1u : 0u;
                if (open) {
                    if (c == kcut) { lo = hi = mid; }
                    else if (c < kcut) lo = mid + 1u; else hi = mid;
                    open = lo < hi;
                }
            }
            const uint32_t thr = lo;
            if (need) {
                uint32_t c_le = 0;
                for (uint32_t e = 0; e < hc; ++e) c_le += hd2[e * 32] <= thr ? 1u : 0u;
                // more than K at or below the threshold: the surplus are ties AT the threshold; those with the largest original index go
                for (; c_le > kcut; --c_le) {
                    uint32_t worst = 0, worst_orig = 0; bool any = false;
                    for (uint32_t e = 0; e < hc; ++e)
                        if (hd2[e * 32] == thr) { const uint32_t og = __ldg(a.m.orig + hpos[e * 32]); if (!any || og > worst_orig) { any = true; worst = e; worst_orig = og; } }
                    hd2[worst * 32] = 0xffffffffu;
                }
                uint32_t j = 0;
                for (uint32_t e = 0; e < hc; ++e) {
                    const uint32_t key = hd2[e * 32];
                    if (key <= thr) { const uint32_t pp = hpos[e * 32]; hd2[j * 32] = key; hpos[j * 32] = pp; ++j; }
                }
                hc = j;
                hbound = __uint_as_float(thr);
            }
            __syncwarp();
        };

        while (remaining) {
            // ---- sub-batch: the queries of the leader's cell row within CG_XSPAN coarse cells of it (normally all 32)
            const int leader = __ffs(remaining) - 1;
            const uint32_t lrow = __shfl_sync(PV_FULL, row, leader);
            const int lx = __shfl_sync(PV_FULL, cxf, leader);
            const bool in = valid && ((remaining >> lane) & 1u) && row == lrow && (uint32_t)(cxf - lx) <= xspan;
            const uint32_t gm = __ballot_sync(PV_FULL, in);
            remaining &= ~gm;
            const v3 aq = in ? q : V3(INFINITY, INFINITY, INFINITY);          // lanes outside the sub-batch accept nothing
            const float slack = MODE ? cg_warp_max(in ? rq : 0.f) + g.margin : slack_fixed;
            const float lox = cg_warp_min(in ? q.x : INFINITY), hix = cg_warp_max(in ? q.x : -INFINITY);
            const float loy = cg_warp_min(in ? q.y : INFINITY), hiy = cg_warp_max(in ? q.y : -INFINITY);
            const float loz = cg_warp_min(in ? q.z : INFINITY), hiz = cg_warp_max(in ? q.z : -INFINITY);
            // ---- block of cells: every cell a photon within maxdist of one of the queries can be in (monotone cell coordinate)
            const int xa = pv_cell_coord(lox - slack, g.origin[0], g.inv_hx, g.dims[0]), xb = pv_cell_coord(hix + slack, g.origin[0], g.inv_hx, g.dims[0]);
            const int y0 = pv_cell_coord(loy - slack, g.origin[1], g.inv_h, g.dims[1]), y1 = pv_cell_coord(hiy + slack, g.origin[1], g.inv_h, g.dims[1]);
            const int z0 = pv_cell_coord(loz - slack, g.origin[2], g.inv_h, g.dims[2]), z1 = pv_cell_coord(hiz + slack, g.origin[2], g.inv_h, g.dims[2]);
            const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
            if (nrows > 32) {                                                  // radius far above the cell size: not this kernel's case
                if (in) { handed_over = true; a.overflow[atomicAdd(a.counters + a.cnt_overflow, 1ull)] = s; }
                continue;
            }
            // lane j < nrows owns row (y0 + j % ny, z0 + j / ny): ONE contiguous photon run [rs, re).  Runs are then ordered by
            // Morton index, i.e. by photon position in the map.
            uint32_t rk = 0xFFFFFFFFu, rs = 0, re = 0;
            if ((int)lane < nrows) {
                rk = pv_morton2((uint32_t)(y0 + (int)lane % ny), (uint32_t)(z0 + (int)lane / ny));
                const uint32_t rowkey = rk << g.xbits;
                rs = __ldg(a.m.cell_start + (rowkey | (uint32_t)xa));
                re = __ldg(a.m.cell_start + (rowkey | (uint32_t)xb) + 1);
            }
            uint32_t rank = 0;
            for (int j = 0; j < nrows; ++j) rank += __shfl_sync(PV_FULL, rk, j) < rk ? 1u : 0u;
            __syncwarp();
            if ((int)lane < nrows) { run_rs[rank] = rs; run_len[rank] = re - rs; }
            __syncwarp();
            uint32_t my_rs = 0, my_len = 0;
            if ((int)lane < nrows) { my_rs = run_rs[lane]; my_len = run_len[lane]; }
            uint32_t inc = my_len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(PV_FULL, inc, o); if (lane >= (uint32_t)o) inc += t; }
            const uint32_t T = __shfl_sync(PV_FULL, inc, 31), E = inc - my_len;
            st_cand += (unsigned long long)T * __popc(gm);

            for (int pass = 0; pass < (HIST ? 2 : 1); ++pass) {
            if (HIST && pass == 1) {
                // ---- between the passes: where does the lane's K-th photon fall?  b* = the first bin at which the running count
                // reaches K; the bins below it hold n_below < K photons, all taken; K - n_below more come from bin b*
                uint32_t nb = 0, bs = CG_NB, nbelow = 0, inb = 0;
                if (in) {
#pragma unroll 1
                    for (int bq = 0; bq < CG_NB; ++bq) { const uint32_t c = hd2[bq * 32]; if (bs == CG_NB && c > 0 && nb + c >= K) { bs = (uint32_t)bq; nbelow = nb; inb = c; } nb += c; }
                    if (nb <= K) { bs = CG_NB; nbelow = nb; inb = 0; }                // K or fewer in the whole trial sphere: every one of them
                    // not this kernel's case: too few inside a trial sphere smaller than maxdist, or a boundary bin that does not fit the list
                    if ((nb < K && rq < a.maxdist) || inb + CG_U > CG_KLIST) { handed_over = true; a.overflow[atomicAdd(a.counters + a.cnt_overflow, 1ull)] = s; r2q = -1.f; }
                    bstar = bs; kcut = bs == CG_NB ? 0u : K - nbelow; hc = 0;
                }
                __syncwarp();
            }
            for (uint32_t cb = 0; cb < T; cb += CG_STAGE) {
                const uint32_t cend = min(T, cb + CG_STAGE), n = cend - cb;
                // ---- stage: generic-proxy reads of the previous round are ordered before the async-proxy writes of this one
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_expect_tx(mbar, n * (iso ? 16u : 32u));
                {
                    const uint32_t lo = max(E, cb), hi = min(E + my_len, cend);
                    if (lo < hi) {
                        tma_bulk_g2s(spos_addr + (lo - cb) * 16u, a.m.pos4 + my_rs + (lo - E), (hi - lo) * 16u, mbar);
                        if (!iso) tma_bulk_g2s(swi_addr + (lo - cb) * 16u, a.m.wi4 + my_rs + (lo - E), (hi - lo) * 16u, mbar);
                    }
                }
                mbar_wait(mbar, phase);
                phase ^= 1u;
                if (lane < CG_U) spos[n + lane] = make_float4(INFINITY, INFINITY, INFINITY, 0.f);      // the unrolled scan may read past n
                __syncwarp();
                if (HIST && pass == 0) {
                    // ---- scan, pass 1 of MODE 2: count the candidates inside the lane's trial sphere by bin of d2
                    for (uint32_t c0 = 0; c0 < n; c0 += CG_U) {
                        float4 p[CG_U];
#pragma unroll
                        for (int u = 0; u < CG_U; ++u) p[u] = spos[c0 + u];
#pragma unroll
                        for (int u = 0; u < CG_U; ++u) {
                            const float dx = p[u].x - aq.x, dy = p[u].y - aq.y, dz = p[u].z - aq.z;
                            const float d2 = dx * dx + dy * dy + dz * dz;
                            if (d2 < r2q) { const uint32_t bq = min((uint32_t)(CG_NB - 1), (uint32_t)(d2 * hscale)); hd2[bq * 32] += 1u; }
                        }
                    }
                    continue;
                }
                if (KNN) {
                    // ---- scan, k-nearest mode: a candidate inside the trial sphere and not beyond the K-th nearest known so far is
                    // appended to the lane's list; a list that could overflow in the next round is cut first
                    for (uint32_t c0 = 0; c0 < n; c0 += CG_U) {
                        if (__reduce_max_sync(PV_FULL, hc) + CG_U > CG_KLIST) knn_cut();
                        float4 p[CG_U];
#pragma unroll
                        for (int u = 0; u < CG_U; ++u) p[u] = spos[c0 + u];
#pragma unroll
                        for (int u = 0; u < CG_U; ++u) {
                            const float dx = p[u].x - aq.x, dy = p[u].y - aq.y, dz = p[u].z - aq.z;
                            const float d2 = dx * dx + dy * dy + dz * dz;
                            if (d2 < r2q && d2 <= hbound) { hd2[hc * 32] = __float_as_uint(d2); hpos[hc * 32] = __float_as_uint(p[u].w); ++hc; }
                        }
                    }
                    continue;                                                      // next stage; the lists are cut and summed at the end of the batch
                }
                // ---- scan (lane == query).  A round of CG_U candidates adds at most CG_U entries to a list, so
                // (CG_CAP - longest list) / CG_U rounds need no check; when a list could overflow, the pending entries are
                // resolved and all lists summed.
                for (uint32_t c0 = 0; c0 < n;) {
                    const uint32_t longest = __reduce_max_sync(PV_FULL, (lp - lidx_addr) >> 7);
                    const uint32_t room = (CG_CAP - longest) / CG_U;               // rounds that cannot overflow any list
                    if (room == 0) { resolve(); sum(); continue; }
                    const uint32_t stop = min(n, c0 + room * CG_U);
                    for (; c0 < stop; c0 += CG_U) {
                        float4 p[CG_U];
#pragma unroll
                        for (int u = 0; u < CG_U; ++u) p[u] = spos[c0 + u];       // broadcast loads first: their latencies overlap
#pragma unroll
                        for (int u = 0; u < CG_U; ++u) {
                            const float dx = p[u].x - aq.x, dy = p[u].y - aq.y, dz = p[u].z - aq.z;
                            const float d2 = dx * dx + dy * dy + dz * dz;          // (p1 - p2).LengthSquared(), geometry.h:116,526
                            if (HIST) {
                                if (d2 < r2q) {
                                    const uint32_t bq = min((uint32_t)(CG_NB - 1), (uint32_t)(d2 * hscale));
                                    if (bq < bstar) {
                                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(lp), "r"(c0 + u) : "memory");
                                        lp += 128u; mx = fmaxf(mx, d2);
                                    } else if (bq == bstar) { hd2[hc * 32] = __float_as_uint(d2); hpos[hc * 32] = __float_as_uint(p[u].w); ++hc; }
                                }
                            } else if (d2 < r2) {
                                asm volatile("st.shared.u32 [%0], %1;" ::"r"(lp), "r"(c0 + u) : "memory");
                                lp += 128u; mx = fmaxf(mx, d2);
                            }
                        }
                    }
                }
                resolve(); sum();                                                  // the stage is about to be overwritten
            }
            }
        }

        if (MODE) {
            knn_cut();
            // a lane short of K photons whose trial sphere was smaller than maxdist has not seen everything: warp-per-step kernel
            if (KNN && valid && !handed_over && hc < K && rq < a.maxdist) { handed_over = true; a.overflow[atomicAdd(a.counters + a.cnt_overflow, 1ull)] = s; }
            if (!valid || handed_over) hc = 0;
            // ---- the heaps, CG_CAP entries at a time, through the same weigh + sum code as the lists
            const uint32_t longest = __reduce_max_sync(PV_FULL, hc);
            for (uint32_t e0 = 0; e0 < longest; e0 += CG_CAP) {
                const uint32_t e1 = min(hc, e0 + CG_CAP);
                for (uint32_t e = e0; e < e1; ++e) {
                    const uint32_t pos = hpos[e * 32];
                    float ph = pc;
                    if (!iso) {
                        const float4 wv = __ldg(a.m.wi4 + pos);
                        const float costheta = -(wv.x * w.x + wv.y * w.y + wv.z * w.z);      // Dot(wi, -w)
                        float rsq;
                        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rsq) : "f"(gg1 - g2 * costheta));
                        ph = pc * rsq * rsq * rsq;
                    }
                    lidx[(e - e0) * 32 + lane] = pos; lw[(e - e0) * 32 + lane] = ph;
                    pad_idx = pos; mx = fmaxf(mx, __uint_as_float(hd2[e * 32]));
                    lp += 128u;
                }
                res = (lp - lidx_addr) >> 7;
                sum();
            }
        }
        // ---- finish: LPhoton's tail (photonvolume.cpp:83-104) per query, one 128-byte row of L_ii per step
        const bool over = !MODE && valid && !handed_over && tot > a.nused;
        {   // needs the k-nearest selection: handed over as ONE run per batch (a single atomic), so that the list keeps the batches'
            // spatial order and the launch that takes it over finds neighbouring steps next to each other
            const uint32_t om = __ballot_sync(PV_FULL, over);
            if (om) {
                unsigned long long at = 0;
                if (lane == 0) at = atomicAdd(a.counters + a.cnt_overflow, (unsigned long long)__popc(om));
                at = __shfl_sync(PV_FULL, at, 0);
                if (over) a.overflow[at + __popc(om & lanemask_lt())] = s;
            }
        }
        const bool done = valid && !handed_over && !over;
        st_found += __reduce_add_sync(PV_FULL, done ? tot : 0u);
        st_heap += __popc(__ballot_sync(PV_FULL, done && tot == a.nused));
        qcnt[lane] = tot; qmx[lane] = mx;
        const uint32_t dmask = __ballot_sync(PV_FULL, done);
        st_lookups += __popc(dmask);
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            const uint32_t ql = grp + 4u * (uint32_t)i;
            if (!((dmask >> ql) & 1u)) continue;
            const uint32_t cq = qcnt[ql];
            const float mq = qmx[ql], dn = qdens[ql];
            const float4 acc = part[ql * 8 + sub];
            float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cq >= 10) {
                const float dV = mq * __fsqrt_rn(mq);
                const float s0 = sg[0] * dn, s1 = sg[1] * dn, s2 = sg[2] * dn, s3 = sg[3] * dn;       // sigma_s(pt)
                const bool any_scale = __ballot_sync(gmask8, s0 != 0.f || s1 != 0.f || s2 != 0.f || s3 != 0.f) != 0;
                if (dV != 0.f && any_scale) {
                    const float f = (float)(4.0 / 3.0 * (double)PV_PI_F * (double)dV);                // 4.0/3.0*M_PI*dV is a double expression
                    out.x = __fdiv_rn(acc.x, s0 * f); out.y = __fdiv_rn(acc.y, s1 * f);
                    if (sub != 7) { out.z = __fdiv_rn(acc.z, s2 * f); out.w = __fdiv_rn(acc.w, s3 * f); }      // bins 30, 31 are padding
                }
            }
            reinterpret_cast<float4 *>(a.lii + (size_t)qstep[ql] * 32)[sub] = out;
        }
        __syncwarp();
    }
    if (a.stats) {
        st_cand = __shfl_sync(PV_FULL, st_cand, 0);
        if (lane == 0) {
            atomicAdd((unsigned long long *)&a.stats->lookups, (unsigned long long)st_lookups);
            atomicAdd((unsigned long long *)&a.stats->photons_found, (unsigned long long)st_found);
            atomicAdd((unsigned long long *)&a.stats->candidates_tested, st_cand);
            atomicAdd((unsigned long long *)&a.stats->heap_lookups, (unsigned long long)st_heap);
        }
    }
}

// ------------------------------------------------------------------ host side
int pvi_cellgather(pv_ctx *ctx, const GatherArgs &ga, const uint32_t **leftover_list, int *leftover_count) {
    const unsigned long long total = ga.total_steps;
    if (total == 0) return PV_OK;
    if (total > 0xFFFFFFF0ull) { ctx->err = "pv_gather: too many march steps in one slice for 32-bit step indices"; return PV_ENOMEM; }
    const GridParams &g = ga.m.g;
    int rc = pv_ensure(ctx, &ctx->cg_sort, &ctx->cg_sort_bytes, (size_t)total * 4 * sizeof(uint32_t) + 256); if (rc) return rc;
    rc = pv_ensure(ctx, &ctx->cg_overflow, &ctx->cg_overflow_bytes, (size_t)total * 2 * sizeof(uint32_t) + 256); if (rc) return rc;
    uint32_t *keys = (uint32_t *)ctx->cg_sort, *vals = keys + total, *keys_tmp = vals + total, *vals_tmp = keys_tmp + total;
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_counters + CG_CNT_BATCH, 0, 4 * sizeof(unsigned long long), ctx->stream));
    cg_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(g, ga.rays, ga.steps, total, keys, vals);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->launches += 1;
    uint32_t *skeys, *svals;
    rc = pvi_sort_pairs_u32(ctx, keys, vals, keys_tmp, vals_tmp, total, std::max(1, g.xbits + 2 * g.yzbits), &skeys, &svals); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->tev[0], ctx->stream));

    CgArgs a;
    a.m = ga.m; a.sc = ga.sc; a.rays = ga.rays; a.steps = ga.steps; a.order = svals; a.total = total; a.maxdist = ga.maxdist; a.nused = ga.nused;
    a.lii = ga.lii; a.overflow = (uint32_t *)ctx->cg_overflow; a.counters = ctx->d_counters; a.stats = ga.stats;
    // Which form?  Search radius within the cells: MODE 0, then MODE 2 over the steps it hands over (more than nused photons in
    // range).  Search radius above the cells (k-nearest regime): MODE 1 while nused fits its per-lane lists, else MODE 2.  What
    // the last launch hands over is left in *leftover_list (count in d_counters[*leftover_count]) for the warp-per-step kernel.
    const bool knn_regime = ga.maxdist > g.h;
    a.knn_rmax = 2.f * g.h - 2.f * g.margin;        // what the 5 x 5 rows (and 5 coarse cells along x) around a query's cell are sure to cover
    a.knn_trial = 1.1f;                              // measured on config 2: 1.05 -> 20.0 ms (3.5 ms of them in the fallback), 1.1 -> 17.7, 1.2 -> 23.1
    if (const char *e = getenv("PV_KNN_TRIAL")) a.knn_trial = std::max(1.0f, std::min(3.0f, (float)atof(e)));       // tuning knob
    uint32_t *ov1 = (uint32_t *)ctx->cg_overflow, *ov2 = ov1 + total;
    static_assert((size_t)CG_WARP_BYTES_KNN * CG_WARPS <= 227 * 1024 && CG_NB <= CG_KLIST, "cellgather: shared memory per CTA");
    auto launch = [&](int mode, const uint32_t *order, const unsigned long long *total_dev, uint32_t cnt_batch, uint32_t *overflow, uint32_t cnt_overflow) -> int {
        void (*kern)(CgArgs) = mode == 0 ? cellgather_kernel<0> : mode == 1 ? cellgather_kernel<1> : cellgather_kernel<2>;
        const size_t smem = (size_t)(mode ? CG_WARP_BYTES_KNN : CG_WARP_BYTES) * CG_WARPS;
        a.order = order; a.total_dev = total_dev; a.cnt_batch = cnt_batch; a.overflow = overflow; a.cnt_overflow = cnt_overflow;
        PV_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        PV_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CG_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        kern<<<ctx->sm_count * per_sm, CG_THREADS, smem, ctx->stream>>>(a);
        ctx->launches += 1;
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        return PV_OK;
    };
    static const bool hist_off = getenv("PV_KNN_NOHIST") != nullptr;          // A/B knob: leave the handed-over steps to the warp-per-step kernel
    if (!knn_regime) {
        rc = launch(0, svals, nullptr, CG_CNT_BATCH, ov1, CG_CNT_OVERFLOW); if (rc) return rc;
        if (!hist_off) {
            rc = launch(2, ov1, ctx->d_counters + CG_CNT_OVERFLOW, CG_CNT_BATCH2, ov2, CG_CNT_OVERFLOW2); if (rc) return rc;
            *leftover_list = ov2; *leftover_count = CG_CNT_OVERFLOW2;
        } else { *leftover_list = ov1; *leftover_count = CG_CNT_OVERFLOW; }
    } else {
        rc = launch(ga.nused <= CG_KMAX ? 1 : 2, svals, nullptr, CG_CNT_BATCH, ov1, CG_CNT_OVERFLOW); if (rc) return rc;
        *leftover_list = ov1; *leftover_count = CG_CNT_OVERFLOW;
    }
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->tev[1], ctx->stream));
    return PV_OK;
}
