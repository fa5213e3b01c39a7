// pv_gather.cu -- the per-camera-ray gather (K5): PhotonVolumeIntegrator::Li
// (integrators/photonvolume.cpp:112-222) with LPhoton (:65-108), KdTree::Lookup
// (core/kdtree.h:150-183) and PhotonProcess (core/photonshooter.h:186-203) fused into one
// persistent sm_100a kernel.
//
// Mapping: ONE WARP PER RAY.
//  * spectral math runs with lane == spectral bin (30 of 32 lanes busy, one register per spectrum);
//  * the per-step scalar work of 32 consecutive march steps (positions, density taps, shadow rays through
//    the LinearBVHNode array, shadow-ray optical depth, Philox draws) runs with lane == step;
//  * a lookup stages the photons {x,y,z,index} of the 3x3x3 cell block around the query into shared memory
//    with TMA bulk copies (cp.async.bulk -> UBLKCP, one copy per contiguous cell row, completion on an
//    mbarrier; the copies of the NEXT march step are issued before the current step's flux sum), then scans
//    them with lane == candidate (two per lane per iteration): ballot/popc compaction of accepted candidates
//    into a per-warp list, radix select (shared-memory histogram) when more than nused are in range;
//  * the flux sum reads each accepted photon's 128-byte alpha line with 8 lanes x float4 (128-bit loads),
//    eight photons per warp iteration.
// Results: neighbour sets are bit-exact (distances use the reference's unfused (dx*dx+dy*dy)+dz*dz on
// identical fp32 positions, ties by photon index); radiance is within 1e-4 relative of the reference
// (summation order and libm differ), see tests/test_gpu_parity.py.
#include <algorithm>
#include <cstdlib>
#include "pv_grid.cuh"
#include "pv_march.cuh"
#include "pv_gather.cuh"

#ifndef GW_WARPS
#define GW_WARPS 4                       // warps per CTA
#endif
#ifndef GW_MIN_CTAS
#define GW_MIN_CTAS 4
#endif
#define GW_THREADS (GW_WARPS * 32)
#ifndef GW_STAGE
#define GW_STAGE 256                     // candidates staged per TMA round (x 16 B)
#endif
#ifndef GW_ALPHA_NOALLOC
#define GW_ALPHA_NOALLOC 1
#endif
#ifndef GW_PRELOAD
#define GW_PRELOAD 1                     // alpha lines of the first 16 photons requested before the phase pass
#endif
#ifndef GW_PF_WI
#define GW_PF_WI 0
#endif
#ifndef GW_STATS
#define GW_STATS 1
#endif
#ifndef GW_LIGHT_REG
#define GW_LIGHT_REG 0                  // measured: the extra live register costs more than the per-step load it saves
#endif
#ifndef GW_PAD16
#define GW_PAD16 0
#endif
#ifndef GW_P2_UNROLL
#define GW_P2_UNROLL 4
#endif
#ifndef GW_RANGES_PRE
// EXPERIMENT, off by default, not yet measured (profiles/r01_v5_gather_regions.md): the nine photon runs of a march step's 3x3
// rows are computed by lanes 0..8 of the ray's warp while 23 lanes idle -- 183 of the kernel's 1112 warp instructions per
// lookup.  With 1 a thread-per-STEP kernel (ranges_kernel) computes them for the whole slice beforehand, all lanes busy, and
// gather_kernel<false> only loads them (80 B per step through ctx->lii, which the ray-parallel form does not use otherwise).
#define GW_RANGES_PRE 0
#endif
#define GW_STR(x) #x
#define GW_UNROLL(n) _Pragma(GW_STR(unroll n))

// per-warp shared memory: candidate list of (d2 bits, photon position) pairs, select histogram, run tables of the
// current batch, mbarrier, TMA staging buffer
// layout per warp: ent[cap] | hist[256] | stats[8] | mbarrier (16 B) | stage[GW_STAGE]
// Only two pointers are kept in registers; everything after the list sits at constant offsets from hdr.
struct WarpBuf { uint2 *ent; unsigned char *hdr; uint32_t cap; uint32_t phase; };
#define WB_HDR_BYTES (1024 + 32 + 16)
__device__ __forceinline__ uint32_t *wb_hist(const WarpBuf &b) { return reinterpret_cast<uint32_t *>(b.hdr); }
__device__ __forceinline__ uint32_t *wb_stats(const WarpBuf &b) { return reinterpret_cast<uint32_t *>(b.hdr + 1024); }
__device__ __forceinline__ float4 *wb_stage(const WarpBuf &b) { return reinterpret_cast<float4 *>(b.hdr + WB_HDR_BYTES); }
enum { ST_LOOKUPS = 0, ST_FOUND, ST_CAND, ST_HEAP };

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(PV_FULL, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(PV_FULL, v, o));
    return v;
}
__device__ __forceinline__ uint32_t wb_mbar(const WarpBuf &b) { return smem_u32(b.hdr + 1024 + 32); }
__device__ __forceinline__ uint32_t wb_stage_addr(const WarpBuf &b) { return smem_u32(b.hdr + WB_HDR_BYTES); }
// Keep the k smallest (d2, original index) entries of ent[0..count); returns the new count (== k) and the
// k-th distance.  count > k on entry, entries hold photon positions.  MSB radix select over the fp32 bit pattern
// (non-negative floats order like unsigned ints), 8-bit digits, histogram in shared memory.  Only reached when
// more than nused photons lie within maxdist, so it is kept out of line (instruction-cache footprint).
__device__ __noinline__ unsigned long long warp_select_k_impl(const uint32_t *__restrict__ orig, uint2 *ent, uint32_t *hist, uint32_t count,
                                                             uint32_t k, uint32_t lane) {
    uint32_t prefix = 0, need = k, m_in_bucket = 0;
    int shift = 24;
    for (int pass = 0; pass < 4; ++pass, shift -= 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
        __syncwarp();
        for (uint32_t e = lane; e < count; e += 32) {
            const uint32_t bits = ent[e].x;
            if (pass == 0 || (bits >> (shift + 8)) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
        }
        __syncwarp();
        uint32_t loc[8], s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { loc[i] = hist[lane * 8 + i]; s += loc[i]; }
        uint32_t inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(PV_FULL, inc, o); if (lane >= o) inc += t; }
        const uint32_t owner = __ffs(__ballot_sync(PV_FULL, inc >= need)) - 1;      // first lane whose inclusive sum reaches need
        uint32_t digit = 0, before = 0, mcount = 0;
        if (lane == owner) {
            uint32_t run = inc - s;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (mcount == 0) {
                    if (run + loc[i] >= need) { digit = lane * 8 + i; before = run; mcount = loc[i]; }
                    else run += loc[i];
                }
            }
        }
        digit = __shfl_sync(PV_FULL, digit, owner); before = __shfl_sync(PV_FULL, before, owner);
        mcount = __shfl_sync(PV_FULL, mcount, owner);
        need -= before; prefix = (prefix << 8) | digit; m_in_bucket = mcount;
        __syncwarp();
        if (m_in_bucket == need) break;                 // the whole bucket is selected: no need to refine further
    }
    if (shift < 0) shift = 0;
    // m_in_bucket > need can only happen after all four passes: ties in d2 straddle the k-th place.
    // Resolve by original photon index (ClosePhoton::operator< tie rule, photonshooter.h:44-47).
    uint32_t tie_idx = 0xFFFFFFFFu;
    if (m_in_bucket != need) {
        uint32_t last = 0; bool first = true;
        for (uint32_t t = 0; t < need; ++t) {
            uint32_t best = 0xFFFFFFFFu;
            for (uint32_t e = lane; e < count; e += 32) {
                if (ent[e].x == prefix) {
                    const uint32_t oi = __ldg(&orig[ent[e].y]);
                    if ((first || oi > last) && oi < best) best = oi;
                }
            }
            last = warp_min_u32(best); first = false;
        }
        tie_idx = last;
    }
    // stable in-place compaction
    uint32_t out = 0, mx = 0;
    for (uint32_t e0 = 0; e0 < count; e0 += 32) {
        const uint32_t e = e0 + lane;
        bool keep = false; uint2 v = make_uint2(0u, 0u);
        if (e < count) {
            v = ent[e];
            const uint32_t hi = v.x >> shift;
            if (hi < prefix) keep = true;
            else if (hi == prefix) {
                if (m_in_bucket == need) keep = true;
                else keep = __ldg(&orig[v.y]) <= tie_idx;
            }
        }
        const uint32_t mask = __ballot_sync(PV_FULL, keep);
        __syncwarp();
        if (keep) { ent[out + __popc(mask & lanemask_lt())] = v; mx = max(mx, v.x); }
        out += __popc(mask);
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(PV_FULL, mx, o));
    return ((unsigned long long)mx << 32) | out;          // (k-th distance bits, new count): no out-pointers, callers keep registers
}
__device__ __forceinline__ uint32_t warp_select_k(const uint32_t *orig, uint2 *ent, uint32_t *hist, uint32_t count, uint32_t k, uint32_t lane,
                                                  float &kth) {
    const unsigned long long r = warp_select_k_impl(orig, ent, hist, count, k, lane);
    kth = __uint_as_float((uint32_t)(r >> 32));
    return (uint32_t)r;
}

// One batch of <= 32 photon runs (lane i holds run [rs, rs+len) of the sorted photon array).  The runs are staged
// back to back into shared memory with TMA bulk copies, GW_STAGE candidates per round, and scanned with
// lane == candidate.  batch_begin() publishes the run tables and issues round 0; batch_finish() waits, scans and
// runs the remaining rounds.  Splitting the two lets the gather issue the NEXT march step's copies before it sums
// the current step's photons, so the copy latency hides behind the alpha loads.
struct Batch { uint32_t rs, len, E, T; };

__device__ __forceinline__ void batch_round_issue(const MapView &m, WarpBuf &b, const Batch &bt, uint32_t cb, uint32_t lane) {
    const uint32_t cend = min(bt.T, cb + GW_STAGE);
    // order the generic-proxy reads of the previous round before the async-proxy writes of this one
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_expect_tx(wb_mbar(b), (cend - cb) * 16u);
    const uint32_t lo = max(bt.E, cb), hi = min(bt.E + bt.len, cend);
    if (lo < hi) tma_bulk_g2s(wb_stage_addr(b) + (lo - cb) * 16u, m.pos4 + bt.rs + (lo - bt.E), (hi - lo) * 16u, wb_mbar(b));
}
__device__ __forceinline__ Batch batch_begin(const MapView &m, WarpBuf &b, uint32_t rs, uint32_t re, uint32_t lane) {
    Batch bt; bt.rs = rs; bt.len = re - rs;
    uint32_t inc = bt.len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(PV_FULL, inc, o); if (lane >= o) inc += t; }
    bt.T = __shfl_sync(PV_FULL, inc, 31);
    bt.E = inc - bt.len;
    if (bt.T == 0) return bt;
    batch_round_issue(m, b, bt, 0, lane);
    return bt;
}
__device__ __forceinline__ void batch_finish(const MapView &m, WarpBuf &b, const Batch &bt, v3 q, float r2, uint32_t k, uint32_t lane,
                                             uint32_t &count, float &boundk, bool &have_k, uint32_t &cand) {
    if (bt.T == 0) return;
    cand += bt.T;
    for (uint32_t cb = 0; cb < bt.T; cb += GW_STAGE) {
        const uint32_t cend = min(bt.T, cb + GW_STAGE);
        if (cb) batch_round_issue(m, b, bt, cb, lane);
        mbar_wait(wb_mbar(b), b.phase);
        b.phase ^= 1u;
        for (uint32_t t0 = cb; t0 < cend; t0 += 64) {         // two candidates per lane per iteration
            const uint32_t ta = t0 + lane, tb = ta + 32;
            float da = INFINITY, db = INFINITY; uint32_t pa = 0, pb = 0;
            if (ta < cend) {
                const float4 pp = wb_stage(b)[ta - cb];
                const float dx = pp.x - q.x, dy = pp.y - q.y, dz = pp.z - q.z;
                da = dx * dx + dy * dy + dz * dz;             // (p1 - p2).LengthSquared(), geometry.h:116,526
                pa = __float_as_uint(pp.w);
            }
            if (tb < cend) {
                const float4 pp = wb_stage(b)[tb - cb];
                const float dx = pp.x - q.x, dy = pp.y - q.y, dz = pp.z - q.z;
                db = dx * dx + dy * dy + dz * dz;
                pb = __float_as_uint(pp.w);
            }
            const bool acca = da < r2 && da <= boundk, accb = db < r2 && db <= boundk;
            const uint32_t ma = __ballot_sync(PV_FULL, acca), mb = __ballot_sync(PV_FULL, accb);
            const uint32_t na = __popc(ma);
            if (acca) b.ent[count + __popc(ma & lanemask_lt())] = make_uint2(__float_as_uint(da), pa);
            if (accb) b.ent[count + na + __popc(mb & lanemask_lt())] = make_uint2(__float_as_uint(db), pb);
#if GW_PF_WI
            // anisotropic phase: the accepted photons' wi will be read by the estimate; pull them towards L1 now
            if (m.need_wi) {
                if (acca) asm volatile("prefetch.global.L1 [%0];" ::"l"(m.wi4 + pa));
                if (accb) asm volatile("prefetch.global.L1 [%0];" ::"l"(m.wi4 + pb));
            }
#endif
            count += na + __popc(mb);
            if (count + 64 > b.cap) {                         // list full: keep the k nearest so far
                __syncwarp();
                count = warp_select_k(m.orig, b.ent, wb_hist(b), count, k, lane, boundk); have_k = true;
            }
        }
    }
    __syncwarp();
}

// Prefetched first batch of a lookup (the 3x3x3 block of the query's cell)
// Only what the common (fixed-radius, in-grid) lookup needs is carried from one march step to the next: the run
// [bt.rs, bt.rs + bt.len) of this lane's row and the batch totals.  The query's cell is recomputed on the rare shell path.
struct Prefetch { bool in_range, issued, fast; Batch bt; };
// clamped cell of a query point: COARSE x cell (the cubic shell geometry of the k-nearest search is in coarse cells), y, z
__device__ __forceinline__ void lookup_cell(const GridParams &g, v3 q, int &cx, int &cy, int &cz) {
    const int ux = (int)floorf((q.x - g.origin[0]) * g.inv_hx), uy = (int)floorf((q.y - g.origin[1]) * g.inv_h),
              uz = (int)floorf((q.z - g.origin[2]) * g.inv_h);
    cx = min(max(ux, 0), g.dims[0] - 1) >> g.xshift; cy = min(max(uy, 0), g.dims[1] - 1); cz = min(max(uz, 0), g.dims[2] - 1);
}

__device__ __forceinline__ bool lookup_in_range(const GridParams &g, v3 q, float r) {
    const float slack = r + g.margin;
    return !(q.x < g.origin[0] - slack || q.x > g.origin[0] + g.dims[0] * g.hx + slack || q.y < g.origin[1] - slack ||
             q.y > g.origin[1] + g.dims[1] * g.h + slack || q.z < g.origin[2] - slack || q.z > g.origin[2] + g.dims[2] * g.h + slack);
}
// The nine rows of the 3x3 (y, z) block around the query, each ONE contiguous photon run: lanes 0..8 load [rs, re).
// A row whose (y, z) slab is at least r away cannot hold a photon with d2 < r2 and is dropped; the others are clipped
// along x to the chord of the search sphere at that row, in fine x cells.  Distances to cell faces are shrunk and the
// chord is widened by the grid margin first, and floor((p - o) * inv_hx) is monotone in p, so the clip is conservative.
__device__ __forceinline__ void lookup_ranges(const MapView &m, v3 q, float r, uint32_t k, uint32_t lane, Prefetch &pf) {
    const GridParams &g = m.g;
    pf.issued = false; pf.fast = false; pf.bt.rs = 0; pf.bt.len = 0; pf.bt.E = 0; pf.bt.T = 0;
    // unclamped cell of the query: inside the grid is the common case and needs no distance test
    const int ux = (int)floorf((q.x - g.origin[0]) * g.inv_hx), uy = (int)floorf((q.y - g.origin[1]) * g.inv_h),
              uz = (int)floorf((q.z - g.origin[2]) * g.inv_h);
    const bool inside = ux >= 0 && ux < g.dims[0] && uy >= 0 && uy < g.dims[1] && uz >= 0 && uz < g.dims[2];
    pf.in_range = m.n != 0 && k != 0 && (inside || lookup_in_range(g, q, r));
    if (!pf.in_range) return;
    const int cx = min(max(ux, 0), g.dims[0] - 1) >> g.xshift, cy = min(max(uy, 0), g.dims[1] - 1), cz = min(max(uz, 0), g.dims[2] - 1);
    pf.fast = inside && r <= g.one_shell_r;              // the 3x3 rows are exhaustive: no shell loop, no radius bookkeeping
    if (lane < 9) {
        const int dy = (int)(lane % 3u) - 1, dz = (int)(lane / 3u) - 1;
        const int y = cy + dy, z = cz + dz;
        if (y >= 0 && y < g.dims[1] && z >= 0 && z < g.dims[2]) {
            // distance from q to the row's slab in y and z (0 for the query's own slab)
            const float ylo = g.origin[1] + cy * g.h, zlo = g.origin[2] + cz * g.h;
            float gy = dy == 0 ? 0.f : (dy < 0 ? q.y - ylo : (ylo + g.h) - q.y);
            float gz = dz == 0 ? 0.f : (dz < 0 ? q.z - zlo : (zlo + g.h) - q.z);
            gy = fmaxf(gy - g.margin, 0.f); gz = fmaxf(gz - g.margin, 0.f);
            const float w2 = r * r - (gy * gy + gz * gz);
            if (w2 > 0.f) {
                const float half = __fsqrt_ru(w2) + g.margin;
                int xa = (int)floorf(((q.x - half) - g.origin[0]) * g.inv_hx), xb = (int)floorf(((q.x + half) - g.origin[0]) * g.inv_hx);
                if (!pf.fast) {                           // shell 1 of the k-nearest search: the coarse cells cx-1 .. cx+1 only
                    xa = max(xa, (cx - 1) << g.xshift); xb = min(xb, ((cx + 2) << g.xshift) - 1);
                }
                xa = max(xa, 0); xb = min(xb, g.dims[0] - 1);
                if (xa <= xb) {
                    const uint32_t rowkey = pv_morton2((uint32_t)y, (uint32_t)z) << g.xbits;
                    pf.bt.rs = __ldg(m.cell_start + (rowkey | (uint32_t)xa));
                    pf.bt.len = __ldg(m.cell_start + (rowkey | (uint32_t)xb) + 1);      // run END until lookup_issue turns it into a length
                }
            }
        }
    }
}
__device__ __forceinline__ void lookup_issue(const MapView &m, WarpBuf &b, uint32_t lane, Prefetch &pf) {
    if (!pf.in_range) return;
    pf.bt = batch_begin(m, b, pf.bt.rs, pf.bt.len, lane);
    pf.issued = true;
}
#if GW_RANGES_PRE
// What lookup_ranges leaves in the Prefetch of lanes 0..8, for one march step: nine [start, end) runs + the two flags.
struct StepRanges { uint2 row[9]; uint32_t flags, pad; };            // 80 bytes; flags: 1 = in_range, 2 = fast
static_assert(sizeof(StepRanges) == 80, "StepRanges");
__device__ __forceinline__ void ranges_load(const StepRanges *sr, uint32_t lane, Prefetch &pf) {
    const uint32_t fl = __ldg(&sr->flags);
    pf.issued = false; pf.in_range = (fl & 1u) != 0; pf.fast = (fl & 2u) != 0; pf.bt.rs = 0; pf.bt.len = 0; pf.bt.E = 0; pf.bt.T = 0;
    if (pf.in_range && lane < 9) { const uint2 v = __ldg(&sr->row[lane]); pf.bt.rs = v.x; pf.bt.len = v.y; }
}
// One THREAD per march step of the slice: the same arithmetic as lookup_ranges, the nine rows in a loop.
__global__ void __launch_bounds__(256) ranges_kernel(MapView m, const pv_ray *__restrict__ rays, const StepRec *__restrict__ steps,
                                                     unsigned long long total, float r, uint32_t k, StepRanges *__restrict__ out) {
    const unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const GridParams &g = m.g;
    const float4 ra = __ldg(reinterpret_cast<const float4 *>(steps + s)), rb = __ldg(reinterpret_cast<const float4 *>(steps + s) + 1);
    const uint32_t ri = __float_as_uint(rb.w);
    const v3 ro = V3(__ldg(&rays[ri].o[0]), __ldg(&rays[ri].o[1]), __ldg(&rays[ri].o[2]));
    const v3 rd = V3(__ldg(&rays[ri].d[0]), __ldg(&rays[ri].d[1]), __ldg(&rays[ri].d[2]));
    const v3 q = ray_at(ro, rd, ra.x);
    StepRanges sr;
#pragma unroll
    for (int i = 0; i < 9; ++i) sr.row[i] = make_uint2(0u, 0u);
    sr.pad = 0u;
    const int ux = (int)floorf((q.x - g.origin[0]) * g.inv_hx), uy = (int)floorf((q.y - g.origin[1]) * g.inv_h),
              uz = (int)floorf((q.z - g.origin[2]) * g.inv_h);
    const bool inside = ux >= 0 && ux < g.dims[0] && uy >= 0 && uy < g.dims[1] && uz >= 0 && uz < g.dims[2];
    const bool in_range = m.n != 0 && k != 0 && (inside || lookup_in_range(g, q, r));
    const bool fast = inside && r <= g.one_shell_r;
    sr.flags = (in_range ? 1u : 0u) | (in_range && fast ? 2u : 0u);
    if (in_range) {
        const int cx = min(max(ux, 0), g.dims[0] - 1) >> g.xshift, cy = min(max(uy, 0), g.dims[1] - 1), cz = min(max(uz, 0), g.dims[2] - 1);
#pragma unroll
        for (int row = 0; row < 9; ++row) {
            const int dy = row % 3 - 1, dz = row / 3 - 1;
            const int y = cy + dy, z = cz + dz;
            if (y >= 0 && y < g.dims[1] && z >= 0 && z < g.dims[2]) {
                const float ylo = g.origin[1] + cy * g.h, zlo = g.origin[2] + cz * g.h;
                float gy = dy == 0 ? 0.f : (dy < 0 ? q.y - ylo : (ylo + g.h) - q.y);
                float gz = dz == 0 ? 0.f : (dz < 0 ? q.z - zlo : (zlo + g.h) - q.z);
                gy = fmaxf(gy - g.margin, 0.f); gz = fmaxf(gz - g.margin, 0.f);
                const float w2 = r * r - (gy * gy + gz * gz);
                if (w2 > 0.f) {
                    const float half = __fsqrt_ru(w2) + g.margin;
                    int xa = (int)floorf(((q.x - half) - g.origin[0]) * g.inv_hx), xb = (int)floorf(((q.x + half) - g.origin[0]) * g.inv_hx);
                    if (!fast) { xa = max(xa, (cx - 1) << g.xshift); xb = min(xb, ((cx + 2) << g.xshift) - 1); }
                    xa = max(xa, 0); xb = min(xb, g.dims[0] - 1);
                    if (xa <= xb) {
                        const uint32_t rowkey = pv_morton2((uint32_t)y, (uint32_t)z) << g.xbits;
                        sr.row[row] = make_uint2(__ldg(m.cell_start + (rowkey | (uint32_t)xa)), __ldg(m.cell_start + (rowkey | (uint32_t)xb) + 1));
                    }
                }
            }
        }
    }
    uint4 *o = reinterpret_cast<uint4 *>(out + s);                       // 80 B = five 128-bit stores
    o[0] = make_uint4(sr.row[0].x, sr.row[0].y, sr.row[1].x, sr.row[1].y); o[1] = make_uint4(sr.row[2].x, sr.row[2].y, sr.row[3].x, sr.row[3].y);
    o[2] = make_uint4(sr.row[4].x, sr.row[4].y, sr.row[5].x, sr.row[5].y); o[3] = make_uint4(sr.row[6].x, sr.row[6].y, sr.row[7].x, sr.row[7].y);
    o[4] = make_uint4(sr.row[8].x, sr.row[8].y, sr.flags, 0u);
}
#endif

// shells s >= 2 of a lookup (k-nearest mode with a sparse neighbourhood): rows on the rim of the (2s+1)^2 square are
// full runs, inner rows contribute their two end cells.  Out of line (the fixed-radius gather never gets here) and
// with everything passed BY VALUE, so the caller's state stays in registers.
struct ShellState { uint32_t count, cand, have_k, phase; float boundk; };
__device__ __noinline__ ShellState lookup_shell(MapView m, WarpBuf b, int s, int cx, int cy, int cz, float qx, float qy, float qz, float r2,
                                                uint32_t k, uint32_t lane, ShellState st) {
    const GridParams &g = m.g;
    const v3 q = V3(qx, qy, qz);
    uint32_t count = st.count, cand = st.cand; float boundk = st.boundk; bool have_k = st.have_k != 0;
    b.phase = st.phase;
    const int side = 2 * s + 1, rows = side * side;
    // x extent of the shell in fine cells: coarse cells cx-s .. cx+s
    const int xs = g.xshift, xlast = g.dims[0] - 1;
    const int x0 = max((cx - s) << xs, 0), x1 = min(((cx + s + 1) << xs) - 1, xlast);
    const float ylo = g.origin[1] + cy * g.h, zlo = g.origin[2] + cz * g.h;
    for (int j0 = 0; j0 < rows; j0 += 32) {
        const int j = j0 + (int)lane;
        uint32_t sa = 0, ea = 0, sb = 0, eb = 0;
        // Nothing farther than the current k-th distance (or r) can enter the result any more: rows of the shell outside
        // that sphere are skipped, the others clipped to its chord -- same conservative margins as lookup_ranges (a tie
        // at exactly the k-th distance lies on the sphere and stays inside the widened chord).
        const float bound2 = fminf(r2, boundk);
        if (j < rows) {
            const int dy = j % side - s, dz = j / side - s;
            const int y = cy + dy, z = cz + dz;
            if (y >= 0 && y < g.dims[1] && z >= 0 && z < g.dims[2]) {
                float gy = dy == 0 ? 0.f : (dy < 0 ? q.y - (ylo + (dy + 1) * g.h) : (ylo + dy * g.h) - q.y);
                float gz = dz == 0 ? 0.f : (dz < 0 ? q.z - (zlo + (dz + 1) * g.h) : (zlo + dz * g.h) - q.z);
                gy = fmaxf(gy - g.margin, 0.f); gz = fmaxf(gz - g.margin, 0.f);
                const float w2 = bound2 - (gy * gy + gz * gz);
                if (w2 > 0.f) {
                    const float half = __fsqrt_ru(w2) + g.margin;
                    const int xa = (int)floorf(((q.x - half) - g.origin[0]) * g.inv_hx), xb = (int)floorf(((q.x + half) - g.origin[0]) * g.inv_hx);
                    const uint32_t rowkey = pv_morton2((uint32_t)y, (uint32_t)z) << g.xbits;
                    if (max(abs(dy), abs(dz)) == s) {
                        const int a0 = max(x0, xa), a1 = min(x1, xb);
                        if (a0 <= a1) { sa = __ldg(m.cell_start + (rowkey | (uint32_t)a0)); ea = __ldg(m.cell_start + (rowkey | (uint32_t)a1) + 1); }
                    } else {
                        // inner rows: the two end (coarse) cells of the shell
                        const int la = max((cx - s) << xs, xa), lb = min(min(((cx - s + 1) << xs) - 1, xlast), xb);
                        const int ra = max((cx + s) << xs, xa), rb = min(min(((cx + s + 1) << xs) - 1, xlast), xb);
                        if (cx - s >= 0 && la <= lb) { sa = __ldg(m.cell_start + (rowkey | (uint32_t)la)); ea = __ldg(m.cell_start + (rowkey | (uint32_t)lb) + 1); }
                        if (ra <= rb) { sb = __ldg(m.cell_start + (rowkey | (uint32_t)ra)); eb = __ldg(m.cell_start + (rowkey | (uint32_t)rb) + 1); }
                    }
                }
            }
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            Batch bt = batch_begin(m, b, half ? sb : sa, half ? eb : ea, lane);
            batch_finish(m, b, bt, q, r2, k, lane, count, boundk, have_k, cand);
        }
    }
    ShellState o; o.count = count; o.cand = cand; o.have_k = have_k ? 1u : 0u; o.phase = b.phase; o.boundk = boundk;
    return o;
}

// KdTree::Lookup + PhotonProcess semantics on the grid: leaves in ent[] the photons with d2 < r2, or, when more
// than k of them exist, the k smallest by (d2, original index).  Returns their number.  `pfp` may carry the first
// batch already in flight (lookup_ranges + lookup_issue called earlier for the same q).
__device__ __forceinline__ uint32_t warp_lookup(const MapView &m, v3 q, float r2, float r, uint32_t k, WarpBuf &b, uint32_t lane,
                                                bool stats, Prefetch *pfp) {
    const GridParams &g = m.g;
    Prefetch pf;
    if (pfp) pf = *pfp; else { lookup_ranges(m, q, r, k, lane, pf); }
    if (!pf.in_range) {                                  // no photon can be within r: still a lookup for the statistics
        if (GW_STATS && stats && lane == 0) wb_stats(b)[ST_LOOKUPS] += 1;
        return 0;
    }
    if (!pf.issued) lookup_issue(m, b, lane, pf);
    uint32_t count = 0, cand = 0;
    bool have_k = false;
    float boundk = INFINITY;
    int cx = 0, cy = 0, cz = 0;
    for (int s = 1;; ++s) {
        if (s == 1) batch_finish(m, b, pf.bt, q, r2, k, lane, count, boundk, have_k, cand);
        else {
            ShellState ss; ss.count = count; ss.cand = cand; ss.have_k = have_k ? 1u : 0u; ss.phase = b.phase; ss.boundk = boundk;
            ss = lookup_shell(m, b, s, cx, cy, cz, q.x, q.y, q.z, r2, k, lane, ss);
            count = ss.count; cand = ss.cand; have_k = ss.have_k != 0; b.phase = ss.phase; boundk = ss.boundk;
        }
        if (count > k) { count = warp_select_k(m.orig, b.ent, wb_hist(b), count, k, lane, boundk); have_k = true; }
        if (s == 1 && pf.fast) break;
        if (s == 1) lookup_cell(g, q, cx, cy, cz);
        // radius up to which the block [c-s, c+s]^3 is guaranteed to contain every photon with d2 < r2
        float gr = INFINITY;
        {
            const float qq[3] = {q.x, q.y, q.z}; const int cc[3] = {cx, cy, cz};
            const int ncoarse[3] = {((g.dims[0] - 1) >> g.xshift) + 1, g.dims[1], g.dims[2]};      // x in coarse cells of width h
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int lo = cc[a] - s, hi = cc[a] + s;
                if (lo > 0) gr = fminf(gr, qq[a] - (g.origin[a] + lo * g.h));
                if (hi < ncoarse[a] - 1) gr = fminf(gr, (g.origin[a] + (hi + 1) * g.h) - qq[a]);
            }
        }
        if (gr == INFINITY) break;                       // the block covers the whole grid
        gr -= g.margin;
        if (gr >= r) break;                              // everything with d2 < r2 has been seen
        if (count == k) {
            if (!have_k) {                               // exactly k so far: the bound is their max distance
                float mx = 0.f;
                for (uint32_t e = lane; e < count; e += 32) mx = fmaxf(mx, __uint_as_float(b.ent[e].x));
                boundk = warp_max(mx); have_k = true;
            }
            if (gr > 0.f && boundk < gr * gr) break;     // strict: an unseen photon cannot even tie
        }
    }
    if (GW_STATS && stats && lane == 0) {
        uint4 *st = reinterpret_cast<uint4 *>(wb_stats(b));           // one 128-bit read-modify-write: {lookups, found, candidates, heap}
        uint4 v = *st;
        v.x += 1; v.y += count; v.z += cand; v.w += count == k ? 1u : 0u;
        *st = v;
    }
    return count;
}

// 128-bit load of a photon's alpha quarter-line that does not allocate in L1: with ~220 KB of the SM's 256 KB carved
// out as shared memory L1 is tiny, and the alpha lines (read once per lookup) would evict everything else.
__device__ __forceinline__ float4 ld_alpha(const float4 *p) {
#if !GW_ALPHA_NOALLOC
    return __ldg(p);
#endif
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// LPhoton tail (photonvolume.cpp:83-104): returns totalFlux[lane] / (4/3 pi r^3 sigma_s[lane]) in lane == bin layout.
// The phase function is not part of any discrete decision, so it uses the fast reciprocal square root
// (<= 2 ulp, far inside the 1e-4 radiance tolerance): PhaseHG = (1-g^2)/(4 pi) * x^-3/2.
__device__ __forceinline__ float warp_estimate(const MapView &m, const DevMedium &med, WarpBuf &b, uint32_t count, v3 w, float dens_pt,
                                               float sig_s_bin, uint32_t lane) {
    if (count < 10) return 0.f;
    const uint32_t grp = lane >> 3, sub = lane & 7;
    // pass 1 (lane == photon): radius of the estimate; the per-photon phase replaces d2 in the entry.
    // The list is padded to a multiple of 8 with zero-weight entries so pass 2 needs no bounds checks
    // (capacity: count <= k and cap >= k + 64).
    float mx = 0.f;
    const float g = med.g, pc = (1.f / (4.f * PV_PI_F)) * (1.f - g * g), gg1 = 1.f + g * g, g2 = 2.f * g;
    const bool iso = g == 0.f;
#if GW_PAD16
    const uint32_t padded = (count + 15u) & ~15u;        // whole 16-photon iterations only: the 8-photon tail (one more exposed L2 round trip) never runs
#else
    const uint32_t padded = (count + 7u) & ~7u;
#endif
    const float4 *a4 = reinterpret_cast<const float4 *>(m.alpha32) + sub;
#if GW_PRELOAD
    // The alpha lines of the first 16 photons (count >= 10, so padded >= 16) are requested BEFORE pass 1: their L2
    // latency overlaps the wi loads and the phase computation instead of following them.
    float4 q0, q1, q2, q3;
    {
        const uint32_t j0 = grp, j1 = grp + 4, j2 = grp + 8, j3 = grp + 12;
        const uint32_t i0 = b.ent[j0].y, i1 = b.ent[j1].y, i2 = j2 < count ? b.ent[j2].y : 0u, i3 = j3 < count ? b.ent[j3].y : 0u;
        q0 = ld_alpha(a4 + (size_t)i0 * 8); q1 = ld_alpha(a4 + (size_t)i1 * 8);
        q2 = ld_alpha(a4 + (size_t)i2 * 8); q3 = ld_alpha(a4 + (size_t)i3 * 8);
    }
#endif
    for (uint32_t e = lane; e < padded; e += 32) {
        if (e < count) {
            const uint2 v = b.ent[e];
            mx = fmaxf(mx, __uint_as_float(v.x));
            float ph = pc;
            if (!iso) {
                const float4 wv = __ldg(m.wi4 + v.y);
                const float costheta = -(wv.x * w.x + wv.y * w.y + wv.z * w.z);      // Dot(wi, -w)
                const float rsq = rsqrtf(gg1 - g2 * costheta);
                ph = pc * rsq * rsq * rsq;
            }
            b.ent[e].x = __float_as_uint(ph);
        } else b.ent[e] = make_uint2(0u, 0u);
    }
    mx = warp_max(mx);
    __syncwarp();
    // pass 2: 8 lanes x float4 per 128-byte alpha line; sixteen photons (four independent 128-bit loads per lane) per
    // iteration so the L2 latency of the alpha lines overlaps instead of adding up
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t e0 = grp;
#if GW_PRELOAD
    {
        const float p0 = __uint_as_float(b.ent[e0].x), p1 = __uint_as_float(b.ent[e0 + 4].x), p2 = __uint_as_float(b.ent[e0 + 8].x),
                    p3 = __uint_as_float(b.ent[e0 + 12].x);
        acc.x = fmaf(q0.x, p0, acc.x); acc.y = fmaf(q0.y, p0, acc.y); acc.z = fmaf(q0.z, p0, acc.z); acc.w = fmaf(q0.w, p0, acc.w);
        acc2.x = fmaf(q1.x, p1, acc2.x); acc2.y = fmaf(q1.y, p1, acc2.y); acc2.z = fmaf(q1.z, p1, acc2.z); acc2.w = fmaf(q1.w, p1, acc2.w);
        acc.x = fmaf(q2.x, p2, acc.x); acc.y = fmaf(q2.y, p2, acc.y); acc.z = fmaf(q2.z, p2, acc.z); acc.w = fmaf(q2.w, p2, acc.w);
        acc2.x = fmaf(q3.x, p3, acc2.x); acc2.y = fmaf(q3.y, p3, acc2.y); acc2.z = fmaf(q3.z, p3, acc2.z); acc2.w = fmaf(q3.w, p3, acc2.w);
        e0 += 16;
    }
#endif
GW_UNROLL(GW_P2_UNROLL)
    for (; e0 + 12 < padded; e0 += 16) {
        const uint2 v0 = b.ent[e0], v1 = b.ent[e0 + 4], v2 = b.ent[e0 + 8], v3_ = b.ent[e0 + 12];
        const float4 x0 = ld_alpha(a4 + (size_t)v0.y * 8), x1 = ld_alpha(a4 + (size_t)v1.y * 8);
        const float4 x2 = ld_alpha(a4 + (size_t)v2.y * 8), x3 = ld_alpha(a4 + (size_t)v3_.y * 8);
        const float p0 = __uint_as_float(v0.x), p1 = __uint_as_float(v1.x), p2 = __uint_as_float(v2.x), p3 = __uint_as_float(v3_.x);
        acc.x = fmaf(x0.x, p0, acc.x); acc.y = fmaf(x0.y, p0, acc.y); acc.z = fmaf(x0.z, p0, acc.z); acc.w = fmaf(x0.w, p0, acc.w);
        acc2.x = fmaf(x1.x, p1, acc2.x); acc2.y = fmaf(x1.y, p1, acc2.y); acc2.z = fmaf(x1.z, p1, acc2.z); acc2.w = fmaf(x1.w, p1, acc2.w);
        acc.x = fmaf(x2.x, p2, acc.x); acc.y = fmaf(x2.y, p2, acc.y); acc.z = fmaf(x2.z, p2, acc.z); acc.w = fmaf(x2.w, p2, acc.w);
        acc2.x = fmaf(x3.x, p3, acc2.x); acc2.y = fmaf(x3.y, p3, acc2.y); acc2.z = fmaf(x3.z, p3, acc2.z); acc2.w = fmaf(x3.w, p3, acc2.w);
    }
#pragma unroll 1
    for (; e0 < padded; e0 += 8) {                            // padded is a multiple of 8: e0 and e0 + 4 are valid
        const uint2 va = b.ent[e0], vb = b.ent[e0 + 4];
        const float4 aa = ld_alpha(a4 + (size_t)va.y * 8), ab = ld_alpha(a4 + (size_t)vb.y * 8);
        const float pha = __uint_as_float(va.x), phb = __uint_as_float(vb.x);
        acc.x = fmaf(aa.x, pha, acc.x); acc.y = fmaf(aa.y, pha, acc.y); acc.z = fmaf(aa.z, pha, acc.z); acc.w = fmaf(aa.w, pha, acc.w);
        acc2.x = fmaf(ab.x, phb, acc2.x); acc2.y = fmaf(ab.y, phb, acc2.y); acc2.z = fmaf(ab.z, phb, acc2.z); acc2.w = fmaf(ab.w, phb, acc2.w);
    }
    acc.x += acc2.x; acc.y += acc2.y; acc.z += acc2.z; acc.w += acc2.w;
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc.x += __shfl_xor_sync(PV_FULL, acc.x, o); acc.y += __shfl_xor_sync(PV_FULL, acc.y, o);
        acc.z += __shfl_xor_sync(PV_FULL, acc.z, o); acc.w += __shfl_xor_sync(PV_FULL, acc.w, o);
    }
    // transpose {8 lanes x 4 bins} -> lane == bin
    const int srcl = (lane >> 2) & 7;
    const float f0 = __shfl_sync(PV_FULL, acc.x, srcl), f1 = __shfl_sync(PV_FULL, acc.y, srcl);
    const float f2 = __shfl_sync(PV_FULL, acc.z, srcl), f3 = __shfl_sync(PV_FULL, acc.w, srcl);
    const int c = lane & 3;
    const float flux = c == 0 ? f0 : (c == 1 ? f1 : (c == 2 ? f2 : f3));
    const float dV = mx * __fsqrt_rn(mx);
    const float scale = sig_s_bin * dens_pt;                            // sigma_s(pt): density * sig_s (or inside ? sig_s : 0)
    const bool any_scale = __ballot_sync(PV_FULL, lane < PV_NSPEC && scale != 0.f) != 0;
    if (dV != 0.f && any_scale) {
        const float f = (float)(4.0 / 3.0 * (double)PV_PI_F * (double)dV);   // 4.0/3.0*M_PI*dV is a double expression
        return __fdiv_rn(flux, scale * f);
    }
    return 0.f;
}

// ------------------------------------------------------------------ kernels
__host__ __device__ __forceinline__ size_t warp_smem_bytes(uint32_t cap) { return (size_t)cap * 8 + WB_HDR_BYTES + (size_t)GW_STAGE * 16; }
__device__ __forceinline__ WarpBuf carve(unsigned char *smem, uint32_t cap, uint32_t warp, uint32_t lane) {
    unsigned char *base = smem + warp_smem_bytes(cap) * warp;
    WarpBuf b;
    b.ent = (uint2 *)base; b.hdr = base + (size_t)cap * 8; b.cap = cap; b.phase = 0;
    if (lane < 8) wb_stats(b)[lane] = 0;
    if (lane == 0) mbar_init(wb_mbar(b), 1);
    __syncwarp();
    return b;
}
__device__ __forceinline__ void flush_stats(pv_gather_stats *gs, const WarpBuf &b, uint32_t rays, uint32_t lane) {
    __syncwarp();
    if (lane == 0 && gs) {
        const uint32_t *st = wb_stats(b);
        atomicAdd((unsigned long long *)&gs->rays, (unsigned long long)rays);
        atomicAdd((unsigned long long *)&gs->lookups, (unsigned long long)st[ST_LOOKUPS]);
        atomicAdd((unsigned long long *)&gs->photons_found, (unsigned long long)st[ST_FOUND]);
        atomicAdd((unsigned long long *)&gs->candidates_tested, (unsigned long long)st[ST_CAND]);
        atomicAdd((unsigned long long *)&gs->heap_lookups, (unsigned long long)st[ST_HEAP]);
    }
}

// bitonic sort of ent[0..n2) by (d2, original index); n2 is a power of two >= count, padding = +inf.
// oidx[] (original photon indices) lives in the unused upper part of the list.
__device__ void warp_sort_entries(const MapView &m, WarpBuf &b, uint32_t count, uint32_t n2, uint32_t *oidx, uint32_t lane) {
    for (uint32_t e = lane; e < n2; e += 32) {
        if (e < count) oidx[e] = __ldg(&m.orig[b.ent[e].y]);
        else { b.ent[e] = make_uint2(__float_as_uint(INFINITY), 0u); oidx[e] = 0xFFFFFFFFu; }
    }
    __syncwarp();
    for (uint32_t size = 2; size <= n2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = lane; t < n2 / 2; t += 32) {
                const uint32_t i = 2 * t - (t & (stride - 1)), j = i + stride;
                const bool up = (i & size) == 0;
                const uint32_t di = b.ent[i].x, dj = b.ent[j].x, oi = oidx[i], oj = oidx[j];
                const bool gt = di > dj || (di == dj && oi > oj);
                if (gt == up) { b.ent[i].x = dj; b.ent[j].x = di; oidx[i] = oj; oidx[j] = oi; }
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(GW_THREADS) knn_kernel(MapView m, const float *__restrict__ pts, uint64_t n, uint32_t k, float r2, uint32_t cap,
                                                        uint32_t *__restrict__ idx, float *__restrict__ d2out, uint32_t *__restrict__ nfound,
                                                        unsigned long long *counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpBuf b = carve(smem, cap, warp, lane);
    const float r = __fsqrt_rn(r2);
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(counter, 1ull);
        q = __shfl_sync(PV_FULL, q, 0);
        if (q >= n) break;
        v3 p = V3(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]);
        uint32_t cnt = warp_lookup(m, p, r2, r, k, b, lane, false, nullptr);
        __syncwarp();
        uint32_t n2 = 2; while (n2 < cnt) n2 <<= 1;
        uint32_t *oidx = reinterpret_cast<uint32_t *>(b.ent + n2);      // the host sizes cap >= 2*n2 (8-byte entries, 4-byte indices)
        warp_sort_entries(m, b, cnt, n2, oidx, lane);
        for (uint32_t e = lane; e < k; e += 32) {
            idx[q * k + e] = e < cnt ? oidx[e] : 0xFFFFFFFFu;
            d2out[q * k + e] = e < cnt ? __uint_as_float(b.ent[e].x) : INFINITY;
        }
        if (lane == 0) nfound[q] = cnt;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(GW_THREADS) lphoton_kernel(MapView m, const DevScene *__restrict__ sc, const float *__restrict__ pts,
                                                            const float *__restrict__ ws, uint64_t n, uint32_t k, float maxdist, uint32_t cap,
                                                            float *__restrict__ L, unsigned long long *counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpBuf b = carve(smem, cap, warp, lane);
    const DevMedium &med = sc->med;
    const float r2 = maxdist * maxdist;
    const float sig_s = lane < PV_NSPEC ? med.sigma_s[lane] : 0.f;
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(counter, 1ull);
        q = __shfl_sync(PV_FULL, q, 0);
        if (q >= n) break;
        v3 p = V3(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]), w = V3(ws[3 * q], ws[3 * q + 1], ws[3 * q + 2]);
        uint32_t cnt = warp_lookup(m, p, r2, maxdist, k, b, lane, false, nullptr);
        __syncwarp();
        float dens = med_density(med, p, nullptr);
        float l = warp_estimate(m, med, b, cnt, w, dens, sig_s, lane);
        if (lane < PV_NSPEC) L[q * PV_NSPEC + lane] = l;
        __syncwarp();
    }
}

// EPhoton (core/photonshooter.cpp:17-35) for every radiance-photon site against ONE surface map: the n_lookup nearest
// photons within max_dist2 (the same lookup as the volume estimate), alpha summed over those arriving on the normal's
// side, divided by path count * md2 * pi where md2 is the search radius^2 as KdTree::Lookup leaves it (the heap's
// largest distance once n_lookup photons were found, else max_dist2).  Accumulates into E32 (32 floats per site).
__global__ void __launch_bounds__(GW_THREADS) ephoton_kernel(MapView m, const float *__restrict__ rp_pos, const float *__restrict__ rp_n,
                                                            const float *__restrict__ rho32, uint64_t n, uint32_t k, float r2, float count,
                                                            uint32_t cap, float *__restrict__ E32, unsigned long long *counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpBuf b = carve(smem, cap, warp, lane);
    const float r = __fsqrt_rn(r2);
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(counter, 1ull);
        q = __shfl_sync(PV_FULL, q, 0);
        if (q >= n) break;
        // `if (!rho_r.IsBlack())` (:371)
        if (__ballot_sync(PV_FULL, lane < PV_NSPEC && rho32[q * 32 + lane] != 0.f) == 0) continue;
        const v3 p = V3(rp_pos[3 * q], rp_pos[3 * q + 1], rp_pos[3 * q + 2]), nn = V3(rp_n[3 * q], rp_n[3 * q + 1], rp_n[3 * q + 2]);
        const uint32_t cnt = warp_lookup(m, p, r2, r, k, b, lane, false, nullptr);
        __syncwarp();
        if (cnt) {
            float mx = 0.f;
            for (uint32_t e = lane; e < cnt; e += 32) mx = fmaxf(mx, __uint_as_float(b.ent[e].x));
            mx = warp_max(mx);
            const float md2 = cnt == k ? mx : r2;
            float acc = 0.f;
            for (uint32_t e = 0; e < cnt; ++e) {
                const uint32_t sp = b.ent[e].y;
                const float4 wv = __ldg(m.wi4 + sp);
                if (nn.x * wv.x + nn.y * wv.y + nn.z * wv.z > 0.f) acc += __ldg(m.alpha32 + (size_t)sp * 32 + lane);
            }
            const float den = (float)((double)(count * md2) * 3.14159265358979323846);          // count * md2 * M_PI
            E32[q * 32 + lane] += __fdiv_rn(acc, den);
        }
        __syncwarp();
    }
}
// PhotonIntegrator's LPhoton, diffuse branch (integrators/photonmap.cpp:62-108, kernel() :57-60): the n_lookup nearest photons of
// the selected surface map within max_dist2, each weighted by the Simpson kernel 3/pi (1 - d2/md2)^2 over nPaths * md2 (md2 = the
// search radius as the lookup leaves it), summed separately for photons arriving on the side of Nf (Lr) and on the other (Lt).
// The caller finishes with L = Lr * rho_r / pi + Lt * rho_t / pi.
__global__ void __launch_bounds__(GW_THREADS) surface_lphoton_kernel(MapView m, const float *__restrict__ pts, const float *__restrict__ nf,
                                                                    uint64_t n, uint32_t k, float r2, float npaths, uint32_t cap,
                                                                    float *__restrict__ Lr, float *__restrict__ Lt, unsigned long long *counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpBuf b = carve(smem, cap, warp, lane);
    const float r = __fsqrt_rn(r2);
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(counter, 1ull);
        q = __shfl_sync(PV_FULL, q, 0);
        if (q >= n) break;
        const v3 p = V3(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]), nn = V3(nf[3 * q], nf[3 * q + 1], nf[3 * q + 2]);
        const uint32_t cnt = warp_lookup(m, p, r2, r, k, b, lane, false, nullptr);
        __syncwarp();
        float accr = 0.f, acct = 0.f;
        if (cnt) {
            float mx = 0.f;
            for (uint32_t e = lane; e < cnt; e += 32) mx = fmaxf(mx, __uint_as_float(b.ent[e].x));
            mx = warp_max(mx);
            const float md2 = cnt == k ? mx : r2;
            const float den = npaths * md2;
            for (uint32_t e = 0; e < cnt; ++e) {
                const uint2 v = b.ent[e];
                const float s = 1.f - __fdiv_rn(__uint_as_float(v.x), md2);
                const float w = __fdiv_rn(((3.f * PV_INV_PI_F) * s) * s, den);
                const float4 wv = __ldg(m.wi4 + v.y);
                const float a = w * __ldg(m.alpha32 + (size_t)v.y * 32 + lane);
                if (nn.x * wv.x + nn.y * wv.y + nn.z * wv.z > 0.f) accr += a; else acct += a;
            }
        }
        if (lane < PV_NSPEC) { Lr[q * PV_NSPEC + lane] = accr; Lt[q * PV_NSPEC + lane] = acct; }
        __syncwarp();
    }
}

// RadiancePhotonProcess + KdTree::Lookup(p, proc, INFINITY) (core/photonshooter.h:54-70, core/kdtree.h:150-183; final gathering,
// integrators/photonmap.cpp:238-243): the NEAREST radiance photon whose normal faces the query normal, no radius limit.  One
// thread per query walks growing shells of grid cells until the best distance found is inside the radius the visited cube
// guarantees.  The grid is the one built over the radiance-photon class: wi4 holds the photon normals.  Ties by photon index.
__device__ __noinline__ uint32_t radiance_nearest(const MapView &m, float px, float py, float pz, float nx, float ny, float nz) {
    const GridParams &g = m.g;
    const v3 p = V3(px, py, pz), nn = V3(nx, ny, nz);
    float best = INFINITY; uint32_t best_i = 0xFFFFFFFFu;
    if (m.n) {
        int cx, cy, cz;
        lookup_cell(g, p, cx, cy, cz);
        const int xs = g.xshift, xlast = g.dims[0] - 1;
        const int ncoarse[3] = {((g.dims[0] - 1) >> xs) + 1, g.dims[1], g.dims[2]};
        const int cc[3] = {cx, cy, cz};
        const float qq[3] = {p.x, p.y, p.z};
        for (int s = 0;; ++s) {
            for (int dz = -s; dz <= s; ++dz) {
                const int z = cz + dz;
                if (z < 0 || z >= g.dims[2]) continue;
                for (int dy = -s; dy <= s; ++dy) {
                    const int y = cy + dy;
                    if (y < 0 || y >= g.dims[1]) continue;
                    const uint32_t rowkey = pv_morton2((uint32_t)y, (uint32_t)z) << g.xbits;
                    const bool rim = max(abs(dy), abs(dz)) == s;
                    // rim rows: the coarse cells cx-s .. cx+s; inner rows: the two end cells of the shell
                    const int nruns = rim || s == 0 ? 1 : 2;
                    for (int h = 0; h < nruns; ++h) {
                        int c0, c1;
                        if (rim) { c0 = cx - s; c1 = cx + s; } else { c0 = c1 = h == 0 ? cx - s : cx + s; }
                        if (c1 < 0 || c0 >= ncoarse[0]) continue;
                        const int a0 = max(c0, 0) << xs, a1 = min(((min(c1, ncoarse[0] - 1) + 1) << xs) - 1, xlast);
                        if (a0 > a1) continue;
                        const uint32_t ps = __ldg(m.cell_start + (rowkey | (uint32_t)a0)), pe = __ldg(m.cell_start + (rowkey | (uint32_t)a1) + 1);
                        for (uint32_t j = ps; j < pe; ++j) {
                            const float4 w = __ldg(m.wi4 + j);
                            if (w.x * nn.x + w.y * nn.y + w.z * nn.z > 0.f) {
                                const float4 pp = __ldg(m.pos4 + j);
                                const float dx = pp.x - p.x, dy2 = pp.y - p.y, dz2 = pp.z - p.z;
                                const float d2 = dx * dx + dy2 * dy2 + dz2 * dz2;
                                if (d2 <= best) {
                                    const uint32_t oi = __ldg(m.orig + j);
                                    if (d2 < best || oi < best_i) { best = d2; best_i = oi; }
                                }
                            }
                        }
                    }
                }
            }
            // radius up to which the cube [c-s, c+s]^3 holds every photon (same rule as warp_lookup)
            float gr = INFINITY;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int lo = cc[a] - s, hi = cc[a] + s;
                if (lo > 0) gr = fminf(gr, qq[a] - (g.origin[a] + lo * g.h));
                if (hi < ncoarse[a] - 1) gr = fminf(gr, (g.origin[a] + (hi + 1) * g.h) - qq[a]);
            }
            if (gr == INFINITY) break;                       // the cube covers the whole grid
            gr -= g.margin;
            if (gr > 0.f && best < gr * gr) break;           // strict: an unseen photon cannot even tie
        }
    }
    return best_i;
}
__global__ void radiance_nearest_kernel(MapView m, const float *__restrict__ pts, const float *__restrict__ nrm, uint64_t n,
                                        uint32_t *__restrict__ idx_out) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    idx_out[q] = radiance_nearest(m, pts[3 * q], pts[3 * q + 1], pts[3 * q + 2], nrm[3 * q], nrm[3 * q + 1], nrm[3 * q + 2]);
}
// One final-gather ray (integrators/photonmap.cpp:231-243 and :278-289): trace it, look the nearest facing radiance photon up at
// the hit (normal = the hit's geometric normal turned towards the ray origin), attenuate its radiance by the transmittance of
// the medium along the ray (renderer->Transmittance with sample == NULL: step 4 * stepSize, one random offset -- here a keyed
// Philox draw per ray).  Lindir = 0 when the ray leaves the scene or no radiance photon faces the hit.
template <bool SPH>
__global__ void final_gather_kernel(MapView m, const DevScene *__restrict__ scp, const pv_ray *__restrict__ rays, uint64_t n, float step,
                                    uint32_t k0, uint32_t k1, uint64_t index_base, const float *__restrict__ Lo32, float *__restrict__ Lindir,
                                    uint32_t *__restrict__ hit_idx) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const DevScene &sc = *scp;
    const pv_ray r = rays[q];
    const v3 o = V3(r.o[0], r.o[1], r.o[2]), d = V3(r.d[0], r.d[1], r.d[2]);
    float thit = r.maxt;
    const int prim = bvh_traverse<false, SPH>(sc, o, d, r.mint, &thit, nullptr);
    uint32_t idx = 0xFFFFFFFFu;
    float s = 0.f;
    if (prim >= 0) {
        v3 hp, nn, dpdu; float eps;
        const float *tv = sc.tri + 9 * (size_t)prim;
        if (!SPH || tv[0] == tv[0]) {                       // shapes/trianglemesh.cpp:160-205 (default uvs), core/diffgeom.cpp:40-55
            const v3 p1 = V3(tv[0], tv[1], tv[2]), p2 = V3(tv[3], tv[4], tv[5]), p3 = V3(tv[6], tv[7], tv[8]);
            const v3 dp1 = p1 - p3, dp2 = p2 - p3;
            dpdu = (dp1 * -1.f - dp2 * -1.f) * 1.f;
            const v3 dpdv = (dp1 * -0.f + dp2 * -1.f) * 1.f;
            nn = vnorm(vcross(dpdu, dpdv));
            hp = ray_at(o, d, thit);
        } else sphere_dg(sc.spheres + (__float_as_uint(tv[0]) & PV_SPHERE_INDEX_MASK), o, d, thit, &hp, &nn, &dpdu, &eps);
        if (vdot(nn, -d) < 0.f) nn = -nn;                   // Faceforward(nGather, -bounceRay.d)
        idx = radiance_nearest(m, hp.x, hp.y, hp.z, nn.x, nn.y, nn.z);
        if (idx != 0xFFFFFFFFu && sc.med.type != PV_MEDIUM_NONE) {
            uint32_t w[4];
            pv_philox4x32_10((uint32_t)(index_base + q), (uint32_t)((index_base + q) >> 32), 0u, PV_RNG_FINAL_GATHER, k0, k1, w);
            s = med_tau_scalar(sc.med, o, d, r.mint, thit, step, pv_u32_to_float(w[0]), nullptr);
        }
    }
    if (hit_idx) hit_idx[q] = idx;
    for (int b = 0; b < PV_NSPEC; ++b)
        Lindir[q * PV_NSPEC + b] = idx == 0xFFFFFFFFu ? 0.f : Lo32[(size_t)idx * 32 + b] * expf(-((sc.med.sigma_a[b] + sc.med.sigma_s[b]) * s));
}

__global__ void gather_lo_kernel(const uint32_t *__restrict__ idx, const float *__restrict__ Lo32, uint64_t n, float *__restrict__ out30) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * PV_NSPEC) return;
    const uint64_t q = i / PV_NSPEC; const uint32_t b = (uint32_t)(i % PV_NSPEC);
    const uint32_t j = idx[q];
    out30[i] = j == 0xFFFFFFFFu ? 0.f : Lo32[(size_t)j * 32 + b];
}

// rp.Lo += INV_PI * rho_r * E (:379)
__global__ void radiance_lo_kernel(const float *__restrict__ rho32, const float *__restrict__ E32, uint64_t n, float *__restrict__ Lo32) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * 32) Lo32[i] = (PV_INV_PI_F * rho32[i]) * E32[i];
}

// RainbowVolume::rainbowReflection (volumes/rainbow.cpp:41-78) for bin `lane`
__device__ __forceinline__ float lerp_or_zero(float x, float x0, float x1, float y0, float y1) {
    if (x < x0 || x1 < x) return 0.f;
    return y0 + __fdiv_rn((x - x0) * (y1 - y0), x1 - x0);
}
__device__ __noinline__ float rainbow_bin(float Ld, v3 w, v3 wi, uint32_t lane) {
    float cosTheta = vdot(wi, -w);
    float theta = 57.2957f * acosf(cosTheta);
    float I = __fdiv_rn(0.5f + 4.5f * powf((float)(0.5 * (double)(1.f + cosTheta)), 8.f), 4.f * PV_PI_F);
    float innerGlow = theta <= 40.4f ? 1.0f : (theta >= 40.45f ? 0.9f : 1.0f + __fdiv_rn((theta - 40.4f) * (0.9f - 1.0f), 40.45f - 40.4f));
    I *= innerGlow;
    float rainbowI = 1.0f;
    const float primaryRainbowI = 0.92f, secondaryRainbowI = (float)(0.42 * (double)0.92f), mistI = 0.08f;
    float lambda = lerp_or_zero(theta, 40.4f, 42.3f, 400.0f, 700.0f);
    if (lambda != 0.f) rainbowI *= primaryRainbowI;
    else {
        lambda = lerp_or_zero(theta, 51.0f, 54.4f, 700.0f, 400.0f);
        if (lambda != 0.f) rainbowI *= secondaryRainbowI;
    }
    if (lambda == 0.f) return Ld * (I * mistI);
    float deltaLambda = 300.f / PV_NSPEC;
    float iwd = __fdiv_rn(lambda - 400.f, deltaLambda);
    int index = (int)iwd;
    float tt = iwd - index;
    float rb = 0.f;
    if ((int)lane == index) rb = Ld * tt;
    if ((int)lane == index + 1 && index + 1 < PV_NSPEC) rb = Ld * (1 - tt);
    return (Ld * mistI + rb * rainbowI) * I;
}

// wo of Light::Sample_L(p, ...) recomputed in the spectral pass (rainbow media only)
__device__ __forceinline__ v3 light_wo(const pv_light &l, v3 p) {
    if (l.type == PV_LIGHT_DISTANT) return V3(l.dir[0], l.dir[1], l.dir[2]);
    return vnorm(V3(l.pos[0], l.pos[1], l.pos[2]) - p);
}

// PRE: the photon estimates L_ii of all steps were computed beforehand by gather_lii_kernel (one warp per STEP); this kernel
// then only runs the recurrence.  The throughput form (PRE == false, one warp does everything for its ray) is what a frame
// of camera rays uses; with a few hundred rays it leaves the GPU idle for as long as one warp needs for its longest ray, so
// small batches (secondary rays of the drop-in) take the step-parallel form.  Same functions on the same inputs: the results
// are bit-identical.
template <bool PRE>
#ifdef GW_MAXNREG
__global__ void __maxnreg__(GW_MAXNREG) gather_kernel(GatherArgs a) {
#else
__global__ void __launch_bounds__(GW_THREADS, GW_MIN_CTAS) gather_kernel(GatherArgs a) {
#endif
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpBuf b = carve(smem, a.cap, warp, lane);
    const DevScene &sc = *a.sc;
    const DevMedium &med = sc.med;
    const bool bin = lane < PV_NSPEC;
    // spectra in lane == bin layout
    const float sig_a = bin ? med.sigma_a[lane] : 0.f, sig_s = bin ? med.sigma_s[lane] : 0.f, le = bin ? med.le[lane] : 0.f;
    const float sig_t = sig_a + sig_s;
    float y_sig_a = 0.f, y_sig_s = 0.f;
    for (int bb = 0; bb < PV_NSPEC; ++bb) { y_sig_a += sc.cie_y[bb] * med.sigma_a[bb]; y_sig_s += sc.cie_y[bb] * med.sigma_s[bb]; }
    const bool y_any = y_sig_a != 0.f || y_sig_s != 0.f;                 // sa.y() != 0 || ss.y() != 0 wherever the density is non-zero
    const float r2 = a.maxdist * a.maxdist;
    const bool rainbow = med.type == PV_MEDIUM_RAINBOW;
    const bool do_lookup = !rainbow && !(a.flags & PV_GATHER_NO_INDIRECT);
    const bool one_light = GW_LIGHT_REG && sc.n_lights == 1;                              // the common case keeps the light's spectrum in a register
    const float I_light0 = (bin && sc.n_lights > 0) ? sc.lights[0].intensity[lane] : 0.f;
    uint32_t nrays = 0;
    for (;;) {
        unsigned long long ri = 0;
        if (lane == 0) ri = atomicAdd(a.counter, 1ull);
        ri = __shfl_sync(PV_FULL, ri, 0);
        if (ri >= a.n) break;
        nrays++;
        const float4 h0 = __ldg(reinterpret_cast<const float4 *>(a.hdr + ri));
        const int nSamples = __float_as_int(h0.z);
        const float step = h0.w;
        const StepRec *recs = a.steps + (((unsigned long long)__float_as_uint(h0.y) << 32) | __float_as_uint(h0.x));
        float Tr = 1.f, Lv = 0.f;
        if (nSamples > 0) {
            const v3 ro = V3(__ldg(&a.rays[ri].o[0]), __ldg(&a.rays[ri].o[1]), __ldg(&a.rays[ri].o[2]));
            const v3 rd = V3(__ldg(&a.rays[ri].d[0]), __ldg(&a.rays[ri].d[1]), __ldg(&a.rays[ri].d[2]));
            bool stop = false;
            for (int c0 = 0; c0 < nSamples && !stop; c0 += 32) {
                // ---------------- lane == step: the march records of the next 32 steps (written by march_steps_kernel)
                float c_t = 0.f, c_tau = 0.f, c_rr = -1.f, c_dens = 0.f, c_sh = 0.f, c_dfac = 0.f;
                int c_ln = 0;
                if (c0 + (int)lane < nSamples) {
                    const float4 *rp = reinterpret_cast<const float4 *>(recs + c0 + lane);
                    const float4 ra = __ldg(rp), rb = __ldg(rp + 1);
                    c_t = ra.x; c_tau = ra.y; c_rr = ra.z; c_dens = ra.w; c_sh = rb.x; c_dfac = rb.y; c_ln = __float_as_int(rb.z);
                }
                // ---------------- lane == bin: the recurrence, one step at a time
                const int nthis = min(32, nSamples - c0);
                Prefetch pf; pf.in_range = false; pf.issued = false;
                for (int i = 0; i < nthis; ++i) {
                    const float s_tau = __shfl_sync(PV_FULL, c_tau, i);
                    const float s_rr = __shfl_sync(PV_FULL, c_rr, i);
                    Tr = expf(-(sig_t * s_tau));                          // Exp(-stepTau): per-step, not cumulative (:155)
                    if (s_rr >= 0.f) {
                        if (s_rr > .5f) {
                            // a prefetched batch must not stay in flight on this warp's mbarrier
                            if (pf.issued && pf.bt.T) { mbar_wait(wb_mbar(b), b.phase); b.phase ^= 1u; }
                            Tr = 0.f; stop = true; break;
                        }
                        Tr = Tr * 2.f;                                    // Tr /= continueProb (0.5): exact
                    }
                    const float s_dens = __shfl_sync(PV_FULL, c_dens, i);
                    const v3 sp = ray_at(ro, rd, __shfl_sync(PV_FULL, c_t, i));
                    const float ss = sig_s * s_dens, sa = sig_a * s_dens;
                    float L_d = 0.f, L_ii = 0.f;
                    const float s_dfac = __shfl_sync(PV_FULL, c_dfac, i);
                    if (s_dfac != 0.f) {
                        const pv_light &lt = sc.lights[one_light ? 0 : __shfl_sync(PV_FULL, c_ln, i)];
                        const float I = one_light ? I_light0 : (bin ? lt.intensity[lane] : 0.f);
                        const float Ld = (I * expf(-(sig_t * __shfl_sync(PV_FULL, c_sh, i)))) * s_dfac;
                        L_d = rainbow ? rainbow_bin(Ld, rd, light_wo(lt, sp), lane) : Ld;
                    }
                    if (PRE) {
                        if (do_lookup) L_ii = __ldg(a.lii + ((size_t)(recs - a.steps) + (size_t)(c0 + i)) * 32 + lane);
                    } else if (do_lookup) {
                        // cell ranges of the NEXT step are requested now and consumed after this step's scan
                        Prefetch nx; nx.in_range = false; nx.issued = false;
                        const bool has_next = i + 1 < nthis;
#if GW_RANGES_PRE
                        const StepRanges *srp = reinterpret_cast<const StepRanges *>(a.lii) + ((size_t)(recs - a.steps) + (size_t)(c0 + i));
                        if (has_next) ranges_load(srp + 1, lane, nx);
                        if (i == 0) ranges_load(srp, lane, pf);             // first step of a chunk: nothing was prefetched for it
                        const uint32_t cnt = warp_lookup(a.m, sp, r2, a.maxdist, a.nused, b, lane, true, &pf);
#else
                        if (has_next) lookup_ranges(a.m, ray_at(ro, rd, __shfl_sync(PV_FULL, c_t, i + 1)), a.maxdist, a.nused, lane, nx);
                        const uint32_t cnt = warp_lookup(a.m, sp, r2, a.maxdist, a.nused, b, lane, true, pf.in_range ? &pf : nullptr);
#endif
                        __syncwarp();
                        if (has_next) lookup_issue(a.m, b, lane, nx);       // next step's TMA copies fly during the flux sum below
                        pf = nx;
                        L_ii = warp_estimate(a.m, med, b, cnt, -rd, s_dens, sig_s, lane);
                        __syncwarp();
                    }
                    // L_i = L_d + (ss/(sa+ss)) * L_ii unless sa.y() == 0 && ss.y() == 0 (photonvolume.cpp:210-213)
                    const float L_i = (s_dens != 0.f && y_any) ? L_d + __fdiv_rn(ss, sa + ss) * L_ii : L_d;
                    Lv = ((sa * (le * s_dens)) * step) + ((ss * L_i) * step) + (Tr * Lv);
                }
            }
        }
        if (bin) { a.L[ri * PV_NSPEC + lane] = Lv; a.T[ri * PV_NSPEC + lane] = Tr; }
    }
    flush_stats(a.stats, b, nrays, lane);
}

// The Lv/Tr recurrence alone (photonvolume.cpp:147-217 without LPhoton): the in-scattered radiance of every march step was
// computed beforehand (gather_lii_kernel / cellgather_kernel, 32 floats per step in a.lii).  Same arithmetic per step as
// gather_kernel -- bit-identical results -- but no shared memory (full occupancy) and the L_ii rows of eight steps are
// requested ahead of the sequential part, so the kernel streams instead of waiting for one row per step.
#define RC_THREADS 256
#define RC_AHEAD 8
__global__ void __launch_bounds__(RC_THREADS) recurrence_kernel(GatherArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const DevScene &sc = *a.sc;
    const DevMedium &med = sc.med;
    const bool bin = lane < PV_NSPEC;
    const float sig_a = bin ? med.sigma_a[lane] : 0.f, sig_s = bin ? med.sigma_s[lane] : 0.f, le = bin ? med.le[lane] : 0.f;
    const float sig_t = sig_a + sig_s;
    float y_sig_a = 0.f, y_sig_s = 0.f;
    for (int bb = 0; bb < PV_NSPEC; ++bb) { y_sig_a += sc.cie_y[bb] * med.sigma_a[bb]; y_sig_s += sc.cie_y[bb] * med.sigma_s[bb]; }
    const bool y_any = y_sig_a != 0.f || y_sig_s != 0.f;
    const bool rainbow = med.type == PV_MEDIUM_RAINBOW;
    const bool do_lookup = !rainbow && !(a.flags & PV_GATHER_NO_INDIRECT);
    const unsigned long long nwarps = ((unsigned long long)gridDim.x * RC_THREADS) >> 5;
    uint32_t nrays = 0;
    for (unsigned long long ri = ((unsigned long long)blockIdx.x * RC_THREADS + threadIdx.x) >> 5; ri < a.n; ri += nwarps) {
        nrays++;
        const float4 h0 = __ldg(reinterpret_cast<const float4 *>(a.hdr + ri));
        const int nSamples = __float_as_int(h0.z);
        const float step = h0.w;
        const unsigned long long first = ((unsigned long long)__float_as_uint(h0.y) << 32) | __float_as_uint(h0.x);
        const StepRec *recs = a.steps + first;
        float Tr = 1.f, Lv = 0.f;
        if (nSamples > 0) {
            const v3 ro = V3(__ldg(&a.rays[ri].o[0]), __ldg(&a.rays[ri].o[1]), __ldg(&a.rays[ri].o[2]));
            const v3 rd = V3(__ldg(&a.rays[ri].d[0]), __ldg(&a.rays[ri].d[1]), __ldg(&a.rays[ri].d[2]));
            bool stop = false;
            for (int c0 = 0; c0 < nSamples && !stop; c0 += 32) {
                float c_t = 0.f, c_tau = 0.f, c_rr = -1.f, c_dens = 0.f, c_sh = 0.f, c_dfac = 0.f;
                int c_ln = 0;
                if (c0 + (int)lane < nSamples) {
                    const float4 *rp = reinterpret_cast<const float4 *>(recs + c0 + lane);
                    const float4 ra = __ldg(rp), rb = __ldg(rp + 1);
                    c_t = ra.x; c_tau = ra.y; c_rr = ra.z; c_dens = ra.w; c_sh = rb.x; c_dfac = rb.y; c_ln = __float_as_int(rb.z);
                }
                const int nthis = min(32, nSamples - c0);
                const float *lrow = a.lii + (first + (unsigned long long)c0) * 32 + lane;
                for (int i0 = 0; i0 < nthis && !stop; i0 += RC_AHEAD) {
                    float li[RC_AHEAD];
#pragma unroll
                    for (int k = 0; k < RC_AHEAD; ++k) li[k] = (do_lookup && i0 + k < nthis) ? __ldg(lrow + (size_t)(i0 + k) * 32) : 0.f;
#pragma unroll
                    for (int k = 0; k < RC_AHEAD; ++k) {
                        const int i = i0 + k;
                        if (i >= nthis) break;
                        const float s_tau = __shfl_sync(PV_FULL, c_tau, i);
                        const float s_rr = __shfl_sync(PV_FULL, c_rr, i);
                        Tr = expf(-(sig_t * s_tau));                          // Exp(-stepTau): per-step, not cumulative (:155)
                        if (s_rr >= 0.f) {
                            if (s_rr > .5f) { Tr = 0.f; stop = true; break; }
                            Tr = Tr * 2.f;                                    // Tr /= continueProb (0.5): exact
                        }
                        const float s_dens = __shfl_sync(PV_FULL, c_dens, i);
                        const float ss = sig_s * s_dens, sa = sig_a * s_dens;
                        float L_d = 0.f;
                        const float s_dfac = __shfl_sync(PV_FULL, c_dfac, i);
                        if (s_dfac != 0.f) {
                            const pv_light &lt = sc.lights[__shfl_sync(PV_FULL, c_ln, i)];
                            const float I = bin ? lt.intensity[lane] : 0.f;
                            const float Ld = (I * expf(-(sig_t * __shfl_sync(PV_FULL, c_sh, i)))) * s_dfac;
                            L_d = rainbow ? rainbow_bin(Ld, rd, light_wo(lt, ray_at(ro, rd, __shfl_sync(PV_FULL, c_t, i))), lane) : Ld;
                        }
                        const float L_ii = li[k];
                        // L_i = L_d + (ss/(sa+ss)) * L_ii unless sa.y() == 0 && ss.y() == 0 (photonvolume.cpp:210-213)
                        const float L_i = (s_dens != 0.f && y_any) ? L_d + __fdiv_rn(ss, sa + ss) * L_ii : L_d;
                        Lv = ((sa * (le * s_dens)) * step) + ((ss * L_i) * step) + (Tr * Lv);
                    }
                }
            }
        }
        if (bin) { a.L[ri * PV_NSPEC + lane] = Lv; a.T[ri * PV_NSPEC + lane] = Tr; }
    }
    if (lane == 0 && a.stats && nrays) atomicAdd((unsigned long long *)&a.stats->rays, (unsigned long long)nrays);
}

// The same recurrence with ONE THREAD PER RAY (frames of camera rays): the 30-bin Lv lives in registers, the medium and light
// spectra in shared memory, a step's StepRec and L_ii row are requested one step ahead.  No shuffles, no idle padding lanes,
// the per-step scalars are plain registers: a quarter of the warp form's instructions.  Same expressions per (step, bin):
// bit-identical results.  Not for rainbow media (they take no photon lookups and never get here).
#define RT_THREADS 128
__global__ void __launch_bounds__(RT_THREADS, 4) recurrence_thread_kernel(GatherArgs a) {
    __shared__ float s_sig_a[PV_NSPEC], s_sig_s[PV_NSPEC], s_le[PV_NSPEC], s_I[PV_MAX_LIGHTS][PV_NSPEC];
    const DevScene &sc = *a.sc;
    const DevMedium &med = sc.med;
    for (int i = threadIdx.x; i < PV_NSPEC; i += RT_THREADS) { s_sig_a[i] = med.sigma_a[i]; s_sig_s[i] = med.sigma_s[i]; s_le[i] = med.le[i]; }
    for (int i = threadIdx.x; i < (int)sc.n_lights * PV_NSPEC; i += RT_THREADS) s_I[i / PV_NSPEC][i % PV_NSPEC] = sc.lights[i / PV_NSPEC].intensity[i % PV_NSPEC];
    __syncthreads();
    const uint64_t ri = (uint64_t)(a.block_order ? a.block_order[blockIdx.x] : blockIdx.x) * RT_THREADS + threadIdx.x;
    if (ri >= a.n) return;
    float y_sig_a = 0.f, y_sig_s = 0.f;
    for (int bb = 0; bb < PV_NSPEC; ++bb) { y_sig_a += sc.cie_y[bb] * med.sigma_a[bb]; y_sig_s += sc.cie_y[bb] * med.sigma_s[bb]; }
    const bool y_any = y_sig_a != 0.f || y_sig_s != 0.f;
    const float4 h0 = __ldg(reinterpret_cast<const float4 *>(a.hdr + ri));
    const int nSamples = __float_as_int(h0.z);
    const float step = h0.w;
    const unsigned long long first = ((unsigned long long)__float_as_uint(h0.y) << 32) | __float_as_uint(h0.x);
    float Lv[PV_NSPEC];
#pragma unroll
    for (int b = 0; b < PV_NSPEC; ++b) Lv[b] = 0.f;
    float last_tau = 0.f, last_rr = -1.f;
    bool stopped = false;
    if (nSamples > 0) {
        const float4 *rp = reinterpret_cast<const float4 *>(a.steps + first);
        const float4 *lp = reinterpret_cast<const float4 *>(a.lii + first * 32);
        float4 na = __ldg(rp), nb = __ldg(rp + 1), nl[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) nl[k] = __ldg(lp + k);
        for (int i = 0; i < nSamples; ++i, rp += 2, lp += 8) {
            const float4 ra = na, rb = nb;
            float l[32];
#pragma unroll
            for (int k = 0; k < 8; ++k) { l[4 * k] = nl[k].x; l[4 * k + 1] = nl[k].y; l[4 * k + 2] = nl[k].z; l[4 * k + 3] = nl[k].w; }
            const float s_rr = ra.z;
            if (s_rr > .5f) { stopped = true; break; }                     // the roulette ends the march (records behind it are dead)
            if (i + 1 < nSamples) {                                          // the next step's records are in flight during this step's bins
                na = __ldg(rp + 2); nb = __ldg(rp + 3);
#pragma unroll
                for (int k = 0; k < 8; ++k) nl[k] = __ldg(lp + 8 + k);
            }
            const float s_tau = ra.y, s_dens = ra.w, s_sh = rb.x, s_dfac = rb.y;
            const bool rr = s_rr >= 0.f, lit = s_dfac != 0.f, mix = s_dens != 0.f && y_any;
            const float *I = s_I[lit ? __float_as_int(rb.z) : 0];
#pragma unroll
            for (int b = 0; b < PV_NSPEC; ++b) {
                const float sig_a = s_sig_a[b], sig_s = s_sig_s[b];
                const float sig_t = sig_a + sig_s;
                float Tr = expf(-(sig_t * s_tau));                           // Exp(-stepTau): per-step, not cumulative (:155)
                if (rr) Tr = Tr * 2.f;                                       // Tr /= continueProb (0.5): exact
                const float ss = sig_s * s_dens, sa = sig_a * s_dens;
                float L_d = 0.f;
                if (lit) L_d = (I[b] * expf(-(sig_t * s_sh))) * s_dfac;
                // L_i = L_d + (ss/(sa+ss)) * L_ii unless sa.y() == 0 && ss.y() == 0 (photonvolume.cpp:210-213)
                const float L_i = mix ? L_d + __fdiv_rn(ss, sa + ss) * l[b] : L_d;
                Lv[b] = ((sa * (s_le[b] * s_dens)) * step) + ((ss * L_i) * step) + (Tr * Lv[b]);
            }
            last_tau = s_tau; last_rr = s_rr;
        }
    }
    float *Lo = a.L + ri * PV_NSPEC, *To = a.T + ri * PV_NSPEC;
#pragma unroll
    for (int b = 0; b < PV_NSPEC; ++b) {
        float Tr = 1.f;
        if (nSamples > 0) {
            Tr = expf(-((s_sig_a[b] + s_sig_s[b]) * last_tau));
            if (last_rr >= 0.f) Tr = Tr * 2.f;
            if (stopped) Tr = 0.f;
        }
        Lo[b] = Lv[b]; To[b] = Tr;
    }
    const uint32_t nr = __popc(__activemask());
    if ((threadIdx.x & 31) == (uint32_t)(__ffs(__activemask()) - 1) && a.stats) atomicAdd((unsigned long long *)&a.stats->rays, (unsigned long long)nr);
}

// One warp per march STEP of the slice: the photon lookup and the radiance estimate of that step (the part of gather_kernel's
// loop body that does not depend on the other steps of the ray).
__global__ void __launch_bounds__(GW_THREADS, GW_MIN_CTAS) gather_lii_kernel(GatherArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpBuf b = carve(smem, a.cap, warp, lane);
    const DevMedium &med = a.sc->med;
    const float sig_s = lane < PV_NSPEC ? med.sigma_s[lane] : 0.f;
    const float r2 = a.maxdist * a.maxdist;
    const unsigned long long nsteps = a.list ? *a.list_count : a.total_steps;
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(a.counter + 3, 1ull);
        g = __shfl_sync(PV_FULL, g, 0);
        if (g >= nsteps) break;
        if (a.list) g = __ldg(a.list + g);
        const float4 *rp = reinterpret_cast<const float4 *>(a.steps + g);
        const float4 ra = __ldg(rp), rb = __ldg(rp + 1);
        const uint32_t ri = __float_as_uint(rb.w);
        const v3 ro = V3(__ldg(&a.rays[ri].o[0]), __ldg(&a.rays[ri].o[1]), __ldg(&a.rays[ri].o[2]));
        const v3 rd = V3(__ldg(&a.rays[ri].d[0]), __ldg(&a.rays[ri].d[1]), __ldg(&a.rays[ri].d[2]));
        const v3 sp = ray_at(ro, rd, ra.x);
        const uint32_t cnt = warp_lookup(a.m, sp, r2, a.maxdist, a.nused, b, lane, true, nullptr);
        __syncwarp();
        const float L_ii = warp_estimate(a.m, med, b, cnt, -rd, ra.w, sig_s, lane);
        __syncwarp();
        a.lii[g * 32 + lane] = L_ii;
    }
    flush_stats(a.stats, b, 0, lane);
}

// ------------------------------------------------------------------ host side
static uint32_t lookup_cap(uint32_t k) {
    uint32_t cap = k + 64u;                          // the list must hold k entries plus one more iteration of 64 candidates
    cap = (cap + 63u) & ~63u;
    // ... and should hold about 2k: a list of k + 64 entries is cut back to the k nearest (warp_select_k: four histogram passes and a
    // compaction over the whole list) after every few dozen accepted candidates once it has filled -- with nused = 500 in a crowded
    // cell that was 25-100 selects per lookup and a third of the kernel's stall samples (profiles/r02_v6_k500_summary.md); with 2k
    // entries a select buys room for k more.  The final list is the same either way (the k nearest in arrival order), so results do
    // not change by a bit.  Bounded by what still lets four CTAs share an SM's shared memory (1088 entries per warp).
    const uint32_t room = (uint32_t)((((227u * 1024u) / 4u - 1024u) / GW_WARPS - WB_HDR_BYTES - GW_STAGE * 16u) / 8u) & ~63u;
    const uint32_t twice = (2u * k + 64u + 63u) & ~63u;
    cap = std::max(cap, std::min(twice, room));
    return std::max<uint32_t>(cap, 256u);
}
MapView pvi_map_view(pv_ctx *ctx) {
    MapView m; m.pos4 = ctx->m_pos4; m.wi4 = ctx->m_wi4; m.alpha32 = ctx->m_alpha32; m.cell_start = ctx->cell_start; m.orig = ctx->m_orig; m.g = ctx->grid;
    m.n = ctx->map_n;
    m.need_wi = ctx->has_scene && ctx->hscene.med.g != 0.f;
    return m;
}
template <typename Kern>
static int launch_cfg(pv_ctx *ctx, Kern kern, uint32_t cap, int *blocks, size_t *smem) {
    *smem = warp_smem_bytes(cap) * GW_WARPS;
    if (*smem > 200 * 1024) { ctx->err = "nused too large for the shared-memory candidate list"; return PV_EINVAL; }
    PV_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem));
    int per_sm = 0;
    PV_CUDA_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GW_THREADS, *smem));
    if (per_sm < 1) per_sm = 1;
    *blocks = ctx->sm_count * per_sm;              // persistent grid: a whole number of CTAs per SM
    return PV_OK;
}

int pvi_knn(pv_ctx *ctx, const float *d_pts, uint64_t n, uint32_t k, float r2, uint32_t *d_idx, float *d_d2, uint32_t *d_nfound) {
    if (!ctx->built) { ctx->err = "pv_knn: photon map not built (call pv_build)"; return PV_ESTATE; }
    if (n == 0 || k == 0) return PV_OK;
    uint32_t n2 = 2; while (n2 < k) n2 <<= 1;
    uint32_t cap = std::max(lookup_cap(k), 2 * n2 + 64);
    cap = (cap + 63u) & ~63u;
    int blocks; size_t smem;
    int rc = launch_cfg(ctx, knn_kernel, cap, &blocks, &smem); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream));
    knn_kernel<<<blocks, GW_THREADS, smem, ctx->stream>>>(pvi_map_view(ctx), d_pts, n, k, r2, cap, d_idx, d_d2, d_nfound, ctx->d_counters);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
int pvi_surface_lphoton(pv_ctx *ctx, const float *d_pts, const float *d_nf, uint64_t n, uint32_t n_lookup, float max_dist2, uint64_t n_paths,
                        float *d_Lr, float *d_Lt) {
    if (n_paths > 0x7fffffffull) { ctx->err = "pv_surface_lphoton: n_paths beyond the reference's int"; return PV_EINVAL; }
    if (!ctx->built || ctx->map_which == PV_MAP_VOLUME || ctx->map_which == PV_MAP_RADIANCE) {
        ctx->err = "pv_surface_lphoton: select a surface photon map first (pv_select_map with PV_MAP_CAUSTIC / INDIRECT / DIRECT)"; return PV_ESTATE;
    }
    if (!(max_dist2 > 0.f) || n_lookup == 0 || n_paths == 0) { ctx->err = "pv_surface_lphoton: n_lookup, max_dist2 and n_paths must be > 0"; return PV_EINVAL; }
    if (n == 0) return PV_OK;
    uint32_t cap = lookup_cap(n_lookup);
    int blocks; size_t smem;
    int rc = launch_cfg(ctx, surface_lphoton_kernel, cap, &blocks, &smem); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream));
    MapView m = pvi_map_view(ctx); m.need_wi = 1;
    surface_lphoton_kernel<<<blocks, GW_THREADS, smem, ctx->stream>>>(m, d_pts, d_nf, n, n_lookup, max_dist2, (float)(int)n_paths, cap, d_Lr, d_Lt,
                                                                    ctx->d_counters);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
int pvi_radiance_nearest(pv_ctx *ctx, const float *d_pts, const float *d_n, uint64_t n, uint32_t *d_idx, float *d_Lo30) {
    if (!ctx->built || ctx->map_which != PV_MAP_RADIANCE) {
        ctx->err = "pv_radiance_nearest: select the radiance-photon map first (pv_select_map with PV_MAP_RADIANCE)"; return PV_ESTATE;
    }
    if (d_Lo30 && !ctx->rad_valid) { ctx->err = "pv_radiance_nearest: radiance not computed (call pv_radiance_photons)"; return PV_ESTATE; }
    if (n == 0) return PV_OK;
    MapView m = pvi_map_view(ctx);
    radiance_nearest_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(m, d_pts, d_n, n, d_idx);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    if (d_Lo30) {
        gather_lo_kernel<<<(unsigned)((n * PV_NSPEC + 255) / 256), 256, 0, ctx->stream>>>(d_idx, ctx->rad_Lo, n, d_Lo30);
        PV_CUDA_CHECK(ctx, cudaGetLastError());
    }
    return PV_OK;
}
int pvi_final_gather(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, float step, uint64_t seed, uint64_t index_base, float *d_Lindir, uint32_t *d_idx) {
    if (!ctx->has_scene) { ctx->err = "pv_final_gather: no scene"; return PV_ESTATE; }
    if (!ctx->built || ctx->map_which != PV_MAP_RADIANCE) {
        ctx->err = "pv_final_gather: select the radiance-photon map first (pv_select_map with PV_MAP_RADIANCE)"; return PV_ESTATE;
    }
    if (!ctx->rad_valid) { ctx->err = "pv_final_gather: radiance not computed (call pv_radiance_photons)"; return PV_ESTATE; }
    if (!(step > 0.f)) { ctx->err = "pv_final_gather: step must be > 0"; return PV_EINVAL; }
    if (n == 0) return PV_OK;
    MapView m = pvi_map_view(ctx);
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (ctx->hscene.n_spheres) final_gather_kernel<true><<<blocks, 128, 0, ctx->stream>>>(m, ctx->dscene, d_rays, n, step, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                                       index_base, ctx->rad_Lo, d_Lindir, d_idx);
    else final_gather_kernel<false><<<blocks, 128, 0, ctx->stream>>>(m, ctx->dscene, d_rays, n, step, (uint32_t)seed, (uint32_t)(seed >> 32), index_base,
                                                                   ctx->rad_Lo, d_Lindir, d_idx);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
int pvi_lphoton(pv_ctx *ctx, const float *d_pts, const float *d_w, uint64_t n, uint32_t nused, float maxdist, float *d_L) {
    if (!ctx->built) { ctx->err = "pv_lphoton: photon map not built (call pv_build)"; return PV_ESTATE; }
    if (ctx->map_which != PV_MAP_VOLUME) { ctx->err = "pv_lphoton: the grid is built over a surface photon map (call pv_build)"; return PV_ESTATE; }
    if (!ctx->has_scene || ctx->hscene.med.type == PV_MEDIUM_NONE) { ctx->err = "pv_lphoton: scene has no medium"; return PV_ESTATE; }
    if (n == 0) return PV_OK;
    uint32_t cap = lookup_cap(nused);
    int blocks; size_t smem;
    int rc = launch_cfg(ctx, lphoton_kernel, cap, &blocks, &smem); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream));
    lphoton_kernel<<<blocks, GW_THREADS, smem, ctx->stream>>>(pvi_map_view(ctx), ctx->dscene, d_pts, d_w, n, nused, maxdist, cap, d_L, ctx->d_counters);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
// ComputeRadianceTask::Run (core/photonshooter.cpp:359-395) for all radiance-photon sites of the context: one grid build and
// one ephoton_kernel launch per surface map, in the reference's summation order direct, indirect, caustic.
int pvi_radiance(pv_ctx *ctx, uint32_t n_lookup, float max_dist2, const uint64_t counts[3]) {
    for (int k = 0; k < 3; ++k) if (counts[k] > 0x7fffffffull) { ctx->err = "pv_radiance_photons: a path count beyond the reference's int"; return PV_EINVAL; }
    PhotonSet &rp = ctx->surf[3];
    ctx->rad_valid = false;
    const uint64_t n = rp.n;
    if (!(max_dist2 > 0.f) || n_lookup == 0) { ctx->err = "pv_radiance_photons: n_lookup and max_dist2 must be > 0"; return PV_EINVAL; }
    if (n == 0) { ctx->rad_valid = true; return PV_OK; }
    if (ctx->rad_Lo_cap < n) {
        if (ctx->rad_Lo) cudaFree(ctx->rad_Lo);
        ctx->rad_Lo = nullptr; ctx->rad_Lo_cap = 0;
        PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->rad_Lo, n * 32 * sizeof(float)));
        ctx->rad_Lo_cap = n;
    }
    int rc = PV_OK;
    float *E32 = ctx->rad_Lo;                             // E accumulates in place (pv_build uses the io buffers), Lo = INV_PI * rho_r * E at the end
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(E32, 0, n * 32 * sizeof(float), ctx->stream));
    const int order[3] = {2, 1, 0};                       // ctx->surf index of direct, indirect, caustic
    const float maxdist = sqrtf(max_dist2);
    for (int k = 0; k < 3 && rc == PV_OK; ++k) {
        PhotonSet &s = ctx->surf[order[k]];
        if (s.n == 0 || counts[k] == 0) continue;
        rc = pvi_build_map(ctx, order[k] + 1, maxdist, n_lookup);      // surface photons are not gated by the medium's extent
        if (rc == PV_OK) {
            uint32_t cap = lookup_cap(n_lookup);
            int blocks; size_t smem;
            rc = launch_cfg(ctx, ephoton_kernel, cap, &blocks, &smem);
            if (rc == PV_OK) {
                cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream);
                MapView m = pvi_map_view(ctx); m.need_wi = 1;
                ephoton_kernel<<<blocks, GW_THREADS, smem, ctx->stream>>>(m, rp.pos, rp.wi, rp.alpha, n, n_lookup, max_dist2, (float)(int)counts[k], cap, E32,
                                                                        ctx->d_counters);
                if (cudaGetLastError() != cudaSuccess) { ctx->err = "ephoton_kernel launch failed"; rc = PV_ECUDA; }
                cudaStreamSynchronize(ctx->stream);
            }
        }
        ctx->built = false;
    }
    if (rc) return rc;
    radiance_lo_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(rp.alpha, E32, n, ctx->rad_Lo);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->rad_valid = true;
    return PV_OK;
}
static unsigned recurrence_blocks(pv_ctx *ctx, uint64_t n) {
    return (unsigned)std::min<uint64_t>((n + RC_THREADS / 32 - 1) / (RC_THREADS / 32), (uint64_t)ctx->sm_count * 8 * 4);
}
// frames: one thread per ray; small batches (fewer rays than resident threads) and rainbow media: one warp per ray
static void launch_recurrence(pv_ctx *ctx, const GatherArgs &a, uint64_t n) {
    if (n >= (uint64_t)ctx->sm_count * 256 && ctx->hscene.med.type != PV_MEDIUM_RAINBOW && ctx->hscene.n_lights <= PV_MAX_LIGHTS)
        recurrence_thread_kernel<<<(unsigned)((n + RT_THREADS - 1) / RT_THREADS), RT_THREADS, 0, ctx->stream>>>(a);
    else
        recurrence_kernel<<<recurrence_blocks(ctx, n), RC_THREADS, 0, ctx->stream>>>(a);
    ctx->launches += 1;
}
// Li for rays [0, n): march records first (pv_march.cu), then the gather kernel.  Rays are taken in slices so that the
// step records of one slice stay within PV_MARCH_MAX_BYTES / the free device memory.
static int gather_slice(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, uint32_t flags, float *d_L, float *d_T) {
    uint64_t total = 0;
    pv_gather_params p = *prm;
    auto halves = [&]() -> int {                           // the records of the slice do not fit: two half slices
        const uint64_t h = n / 2;
        int r = gather_slice(ctx, d_rays, h, prm, flags, d_L, d_T); if (r) return r;
        pv_gather_params p2 = *prm; p2.ray_index_base = prm->ray_index_base + h;
        return gather_slice(ctx, d_rays + h, n - h, &p2, flags, d_L + h * PV_NSPEC, d_T + h * PV_NSPEC);
    };
    int rc = pvi_march(ctx, d_rays, n, &p, flags, &total);
    if (rc == PV_ENOMEM && n > 4096) return halves();
    if (rc == PV_ENOMEM) ctx->err = "pv_gather: out of device memory for the march records";
    if (rc) return rc;
    GatherArgs a;
    a.m = pvi_map_view(ctx); a.sc = ctx->dscene; a.rays = d_rays; a.hdr = (const RayHdr *)ctx->march_hdr; a.steps = (const StepRec *)ctx->march_steps;
    a.n = n; a.maxdist = prm->maxdist;
    a.nused = prm->nused; a.flags = flags; a.cap = lookup_cap(std::max<uint32_t>(prm->nused, 1u));
    a.L = d_L; a.T = d_T; a.stats = ctx->d_stats; a.counter = ctx->d_counters;
    a.lii = nullptr; a.total_steps = total; a.list = nullptr; a.list_count = nullptr;
    static_assert(RT_THREADS == 128, "recurrence_thread_kernel shares the march kernels' blocks of 128 rays");
    a.block_order = ctx->march_blk ? (const uint32_t *)ctx->march_blk + (n + 127) / 128 : nullptr;
    int blocks; size_t smem;
    // Which schedule for the lookups?  (results: cell-batched sums each step's photons in photon order; the two warp forms are
    // bit-identical to each other and agree with it to rounding)
    //   cell-batched   steps sorted by photon-grid cell, one warp per batch of 32 (pv_cellgather.cu): the default whenever the
    //                  search radius fits the cells (fixed-radius regime); steps with more than nused photons in range fall
    //                  through to the warp-per-step kernel;
    //   step-parallel  one warp per march step (k-nearest regime, small batches);
    //   ray-parallel   one warp per ray, lookups fused with the recurrence (k-nearest regime, frames).
    // PV_GATHER_CELL_BATCHED / PV_GATHER_STEP_PARALLEL / PV_GATHER_RAY_PARALLEL in params->flags force one.
    const bool lookups = !(flags & PV_GATHER_NO_INDIRECT) && total > 0;
    static const bool legacy_default = getenv("PV_GATHER_LEGACY") != nullptr;               // A/B knob: the round-1 schedules
    // (k-nearest regime, maxdist above the cell size: the cell-batched kernel's k-nearest mode while nused fits its per-lane heap)
    static const bool knn_legacy = getenv("PV_KNN_LEGACY") != nullptr;                       // A/B knob: warp-per-ray / warp-per-step k-nearest search
    // (nused > 64 in the k-nearest regime stays with the warp-per-ray search: the radius-histogram mode of the batched kernel handles
    // any nused, but on pinkfloyd.pbrt -- 5 M photons crowded into a spot beam inside a 15^3 medium, cells at the 256-per-axis limit --
    // its two passes over blocks of tens of thousands of candidates were an order of magnitude SLOWER than the 35 s of that search)
    bool cell = lookups && !legacy_default && a.m.n > 0 && (prm->maxdist <= ctx->grid.h || (!knn_legacy && prm->nused <= 64 && prm->nused >= 1));
    bool step_parallel = lookups && n < (uint64_t)ctx->sm_count * 16 * 2 && total <= (4ull << 20);
    if (flags & PV_GATHER_STEP_PARALLEL) { step_parallel = lookups; cell = false; }
    if (flags & PV_GATHER_RAY_PARALLEL) { step_parallel = false; cell = false; }
    if (flags & PV_GATHER_CELL_BATCHED) cell = lookups && a.m.n > 0;
    if (cell) step_parallel = false;
    if (step_parallel || cell) {
        rc = pv_ensure(ctx, &ctx->lii, &ctx->lii_bytes, (size_t)total * 32 * sizeof(float));
        if (rc == PV_ENOMEM && n > 4096) return halves();
        if (rc) return rc;
        a.lii = (float *)ctx->lii;
    }
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
#if GW_RANGES_PRE
    if (!step_parallel && !cell && lookups) {               // inside the ev0..ev1 bracket: the kernel time reported stays comparable
        rc = pv_ensure(ctx, &ctx->lii, &ctx->lii_bytes, (size_t)total * sizeof(StepRanges)); if (rc) return rc;
        a.lii = (float *)ctx->lii;
        ranges_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(a.m, d_rays, a.steps, total, prm->maxdist, a.nused,
                                                                                 (StepRanges *)ctx->lii);
        PV_CUDA_CHECK(ctx, cudaGetLastError());
    }
#endif
    if (cell) {
        const uint32_t *left = nullptr; int left_cnt = CG_CNT_OVERFLOW;
        rc = pvi_cellgather(ctx, a, &left, &left_cnt);      // records tev[0] (steps sorted) and tev[1] (cellgather kernels done)
        if (rc == PV_ENOMEM && n > 4096) return halves();
        if (rc) return rc;
        // the steps they left over: warp-per-step kernel over that list
        GatherArgs o = a;
        o.list = left; o.list_count = ctx->d_counters + left_cnt;
        rc = launch_cfg(ctx, gather_lii_kernel, a.cap, &blocks, &smem); if (rc) return rc;
        gather_lii_kernel<<<blocks, GW_THREADS, smem, ctx->stream>>>(o);
        ctx->launches += 1;
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->tev[2], ctx->stream));
        launch_recurrence(ctx, a, n);
    } else if (step_parallel) {
        rc = launch_cfg(ctx, gather_lii_kernel, a.cap, &blocks, &smem); if (rc) return rc;
        gather_lii_kernel<<<blocks, GW_THREADS, smem, ctx->stream>>>(a);
        ctx->launches += 1;
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        launch_recurrence(ctx, a, n);
    } else {
        rc = launch_cfg(ctx, gather_kernel<false>, a.cap, &blocks, &smem); if (rc) return rc;
        gather_kernel<false><<<blocks, GW_THREADS, smem, ctx->stream>>>(a);
        ctx->launches += 1;
    }
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    // device times of this slice (the next slice starts with a stream synchronisation anyway)
    PV_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms += ms;
    cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3); ctx->last_march_ms += ms;
    if (cell) {
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->tev[0]); ctx->phase_ms[0] += ms;
        cudaEventElapsedTime(&ms, ctx->tev[0], ctx->tev[1]); ctx->phase_ms[1] += ms;
        cudaEventElapsedTime(&ms, ctx->tev[1], ctx->tev[2]); ctx->phase_ms[2] += ms;
        cudaEventElapsedTime(&ms, ctx->tev[2], ctx->ev1); ctx->phase_ms[3] += ms;
    }
    return PV_OK;
}
int pvi_gather(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, float *d_L, float *d_T) {
    if (!ctx->has_scene) { ctx->err = "pv_gather: no scene"; return PV_ESTATE; }
    bool need_map = !(prm->flags & PV_GATHER_NO_INDIRECT) && ctx->hscene.med.type != PV_MEDIUM_RAINBOW && ctx->hscene.med.type != PV_MEDIUM_NONE;
    if (need_map && !ctx->built) { ctx->err = "pv_gather: photon map not built (call pv_build)"; return PV_ESTATE; }
    if (need_map && ctx->map_which != PV_MAP_VOLUME) { ctx->err = "pv_gather: the grid is built over a surface photon map (call pv_build)"; return PV_ESTATE; }
    if (!(prm->stepsize > 0.f)) { ctx->err = "pv_gather: stepsize must be > 0"; return PV_EINVAL; }
    if (need_map && !(prm->maxdist > 0.f)) { ctx->err = "pv_gather: maxdist must be > 0"; return PV_EINVAL; }
    if (need_map && (prm->nused < 1 || warp_smem_bytes(lookup_cap(prm->nused)) * GW_WARPS > 200 * 1024)) {
        ctx->err = "pv_gather: nused must be >= 1 and small enough for the shared-memory candidate list"; return PV_EINVAL;
    }
    ctx->last_ms = 0.f; ctx->last_march_ms = 0.f;
    for (int i = 0; i < 4; ++i) ctx->phase_ms[i] = 0.f;
    if (n == 0) return PV_OK;
    const uint32_t flags = prm->flags | (need_map ? 0u : PV_GATHER_NO_INDIRECT);
    const uint64_t slice = PV_GATHER_SLICE_RAYS;
    for (uint64_t r0 = 0; r0 < n; r0 += slice) {
        const uint64_t nr = std::min<uint64_t>(slice, n - r0);
        pv_gather_params p = *prm; p.ray_index_base = prm->ray_index_base + r0;
        int rc = gather_slice(ctx, d_rays + r0, nr, &p, flags, d_L + r0 * PV_NSPEC, d_T + r0 * PV_NSPEC);
        if (rc) return rc;
    }
    return PV_OK;
}
