// pv_build.cu -- the photon map: replaces KdTree<Photon>::KdTree / recursiveBuild
// (core/kdtree.h:99-147, single-threaded nth_element) with a GPU-built uniform grid:
//   1. bounding box of the photons                      (block reduce + ordered-int atomics)
//   2. cell key per photon: (morton2(cy,cz) << xbits) | cx
//   3. stable LSD radix sort of (key, photon index) pairs, 8-bit digits, ranks resolved in
//      shared memory (per-warp match-any multisplit + per-block digit offsets)
//   4. gather the 160-byte photon records into sorted order (pos4 | wi4 | alpha32)
//   5. cell_start[key] = lower_bound(sorted keys, key)
// Everything is HBM-bound streaming work; algorithmic bytes per photon are listed in DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "pv_grid.cuh"

// ------------------------------------------------------------------ exclusive scan (u32)
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(uint32_t *data, uint64_t n, uint32_t *tile_sums) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? data[base + i] : 0u; sum += v[i]; }
    uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(PV_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(PV_FULL, wi, o); if (lane >= o) wi += t; }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + inc - sum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) data[base + i] = run; run += v[i]; }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t *data, uint64_t n, const uint32_t *tile_offsets) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t off = tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) data[base + i] += off;
}
// scratch must hold ceil(n/TILE) + ceil(that/TILE) + ... + 1 u32
static int exclusive_scan_u32(pv_ctx *ctx, uint32_t *data, uint64_t n, uint32_t *scratch) {
    if (n == 0) return PV_OK;
    uint32_t tiles = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    scan_tiles_kernel<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(data, n, scratch);
    ctx->launches += 1;
    if (tiles > 1) {
        int rc = exclusive_scan_u32(ctx, scratch, tiles, scratch + tiles);
        if (rc) return rc;
        scan_add_kernel<<<tiles, SCAN_THREADS, 0, ctx->stream>>>(data, n, scratch);
        ctx->launches += 1;
    }
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
static size_t scan_scratch_words(uint64_t n) {
    size_t w = 0;
    while (n > 1) { n = (n + SCAN_TILE - 1) / SCAN_TILE; w += n; if (n == 1) break; }
    return w + 4;
}

// ------------------------------------------------------------------ LSD radix sort of (key, value) pairs
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(const K *__restrict__ keys, uint64_t n, int shift, uint32_t *__restrict__ hist,
                                                             uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];      // digit-major for one global scan
}

// Stable scatter.  Each warp owns a contiguous run of 32*RS_ITEMS keys; ranks come from match-any
// peers plus a running per-warp digit counter in shared memory, then per-block digit offsets.
template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const K *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
                                                               int shift, const uint32_t *__restrict__ offsets, uint32_t nblocks,
                                                               K *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    __shared__ uint32_t warp_hist[RS_WARPS][256];
    __shared__ uint32_t digit_base[256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
    digit_base[threadIdx.x] = offsets[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();
    uint64_t wbase = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)warp * (32 * RS_ITEMS);
    K k[RS_ITEMS]; uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        bool valid = i < n;
        k[r] = valid ? keys[i] : (K)0;
        uint32_t d = valid ? ((uint32_t)(k[r] >> shift) & 255u) : 256u;      // 256: invalid lanes form their own peer group
        uint32_t peers = __match_any_sync(PV_FULL, d);
        uint32_t pre = valid ? warp_hist[warp][d] : 0u;
        __syncwarp();
        if (valid && (peers & lanemask_lt()) == 0) warp_hist[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + __popc(peers & lanemask_lt());
    }
    __syncthreads();
    {   // exclusive scan over the warps of this block, per digit (thread == digit)
        uint32_t off = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t t = warp_hist[w][threadIdx.x]; warp_hist[w][threadIdx.x] = off; off += t; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint64_t i = wbase + (uint64_t)r * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
            uint32_t dst = digit_base[d] + warp_hist[warp][d] + rank[r];
            keys_out[dst] = k[r];
            vals_out[dst] = vals[i];
        }
    }
}

template <typename K>
static int sort_pairs(pv_ctx *ctx, K *keys, uint32_t *vals, K *keys_tmp, uint32_t *vals_tmp, uint64_t n, int key_bits, K **keys_out,
                      uint32_t **vals_out) {
    *keys_out = keys; *vals_out = vals;
    if (n == 0) return PV_OK;
    uint32_t nblocks = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint64_t hist_n = (uint64_t)256 * nblocks;
    size_t need = (hist_n + scan_scratch_words(hist_n)) * sizeof(uint32_t);
    // the histogram lives in its own context-owned buffer (ctx->scratch may hold the key/value arrays): no allocation, no
    // synchronisation per sort
    int rc = pv_ensure(ctx, &ctx->sort_hist, &ctx->sort_hist_bytes, need); if (rc) return rc;
    uint32_t *hist = (uint32_t *)ctx->sort_hist, *sscratch = hist + hist_n;
    int passes = (key_bits + 7) / 8;
    K *kin = keys, *kout = keys_tmp; uint32_t *vin = vals, *vout = vals_tmp;
    for (int p = 0; p < passes && rc == PV_OK; ++p) {
        int shift = 8 * p;
        rs_count_kernel<K><<<nblocks, RS_THREADS, 0, ctx->stream>>>(kin, n, shift, hist, nblocks);
        rc = exclusive_scan_u32(ctx, hist, hist_n, sscratch);
        if (rc) break;
        rs_scatter_kernel<K><<<nblocks, RS_THREADS, 0, ctx->stream>>>(kin, vin, n, shift, hist, nblocks, kout, vout);
        ctx->launches += 2;                                 // + rs_count_kernel above
        std::swap(kin, kout); std::swap(vin, vout);
    }
    if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    *keys_out = kin; *vals_out = vin;
    return PV_OK;
}
int pvi_sort_pairs_u32(pv_ctx *ctx, uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int key_bits,
                       uint32_t **keys_out, uint32_t **vals_out) {
    return sort_pairs<uint32_t>(ctx, keys, vals, keys_tmp, vals_tmp, n, key_bits, keys_out, vals_out);
}
int pvi_sort_pairs_u64(pv_ctx *ctx, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int key_bits,
                       uint64_t **keys_out, uint32_t **vals_out) {
    return sort_pairs<uint64_t>(ctx, keys, vals, keys_tmp, vals_tmp, n, key_bits, keys_out, vals_out);
}

// ------------------------------------------------------------------ bounding box
__device__ __forceinline__ int float_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__host__ __device__ __forceinline__ float ordered_float(int i) {
    int j = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(j);
#else
    float f; memcpy(&f, &j, 4); return f;
#endif
}
__global__ void bbox_kernel(const float *__restrict__ pos, uint64_t n, int *__restrict__ out /* min xyz, max xyz (ordered ints) */) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { float v = pos[3 * i + a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(PV_FULL, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(PV_FULL, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&out[a], float_ordered(mn[a])); atomicMax(&out[3 + a], float_ordered(mx[a])); }
    }
}

// ------------------------------------------------------------------ keys, record gather, cell table
__global__ void keys_kernel(const float *__restrict__ pos, uint64_t n, GridParams g, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = pv_cell_coord(pos[3 * i + 0], g.origin[0], g.inv_hx, g.dims[0]);
    int cy = pv_cell_coord(pos[3 * i + 1], g.origin[1], g.inv_h, g.dims[1]);
    int cz = pv_cell_coord(pos[3 * i + 2], g.origin[2], g.inv_h, g.dims[2]);
    keys[i] = pv_cell_key(g.xbits, cx, cy, cz);
    vals[i] = (uint32_t)i;
}
// One warp moves four photons per iteration: 8 lanes x float4 per 128-byte alpha line (read and write),
// the first lane of each group also moves pos/wi into the float4 planes.
__global__ void gather_records_kernel(const uint32_t *__restrict__ order, uint64_t n, const float *__restrict__ pos,
                                      const float *__restrict__ wi, const float *__restrict__ alpha, float4 *__restrict__ pos4,
                                      float4 *__restrict__ wi4, float *__restrict__ alpha32, uint32_t *__restrict__ orig, const DevScene *__restrict__ sc) {
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    // Homogeneous media gate the phase function by extent.Inside(photon position) (volumes/homogeneous.h:74-77):
    // a photon outside contributes alpha * 0.  Fold that factor into the map's copy of alpha.
    const bool gate = sc && med_is_homog(sc->med);
    for (uint64_t j0 = warp * 4; j0 < n; j0 += nwarps * 4) {
        uint64_t j = j0 + grp;
        if (j >= n) continue;
        uint32_t src = order[j];
        float px = pos[(uint64_t)src * 3], py = pos[(uint64_t)src * 3 + 1], pz = pos[(uint64_t)src * 3 + 2];
        float4 a = *(reinterpret_cast<const float4 *>(alpha + (uint64_t)src * 32) + sub);
        if (gate && !bbox_inside(sc->med.p0, sc->med.p1, med_to_volume_p(sc->med, V3(px, py, pz)))) a = make_float4(0.f, 0.f, 0.f, 0.f);
        *(reinterpret_cast<float4 *>(alpha32 + j * 32) + sub) = a;
        if (sub == 0) {
            pos4[j] = make_float4(px, py, pz, __uint_as_float((uint32_t)j));
            orig[j] = src;
            wi4[j] = make_float4(wi[(uint64_t)src * 3], wi[(uint64_t)src * 3 + 1], wi[(uint64_t)src * 3 + 2], 0.f);
        }
    }
}
// cell_start[c] = number of photons whose key is < c, for c in [0, table_size]: ONE pass over the sorted keys -- thread i looks
// at keys i - 1 and i and, where they differ, i is written to every cell in (key[i-1], key[i]] (cells in between are empty and start
// where the next occupied one does); the last thread closes the table.  Short runs of empty cells are filled by the thread itself;
// a long one (the padding of the Morton-ordered key space, the empty space around a crowded map: millions of cells) is filled by
// its whole warp, 32 cells per step -- one thread doing that alone took 0.1-0.3 s per build on a 2^24-cell table.
__global__ void __launch_bounds__(256) cell_start_kernel(const uint32_t *__restrict__ sorted_keys, uint64_t n, uint32_t table_size,
                                                         uint32_t *__restrict__ cell_start) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    long long lo = 0, hi = -1;                                                // cells lo .. hi start at i
    if (i <= n) {
        hi = i < n ? (long long)sorted_keys[i] : (long long)table_size;
        lo = i == 0 ? 0ll : (long long)sorted_keys[i - 1] + 1;
    }
    const bool big = hi - lo >= 64;
    if (!big) for (long long c = lo; c <= hi; ++c) cell_start[c] = (uint32_t)i;
    uint32_t todo = __ballot_sync(0xffffffffu, big);
    while (todo) {
        const int src = __ffs(todo) - 1; todo &= todo - 1;
        const long long l = __shfl_sync(0xffffffffu, lo, src), h = __shfl_sync(0xffffffffu, hi, src);
        const uint32_t v = (uint32_t)__shfl_sync(0xffffffffu, (unsigned long long)i, src);
        for (long long c = l + lane; c <= h; c += 32) cell_start[c] = v;
    }
}

static int ceil_log2(int v) { int b = 0; while ((1 << b) < v) ++b; return b; }

// sum over cells of (photons in the cell)^2: divided by n it is the occupancy of the cell a photon picked at random sits in, i.e.
// what a lookup near the photons has to scan per cell -- the mean occupancy says nothing about a map whose photons crowd along a beam
__global__ void __launch_bounds__(256) occupancy_kernel(const uint32_t *__restrict__ cell_start, uint32_t table_size, unsigned long long *out) {
    unsigned long long s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < table_size; i += gridDim.x * blockDim.x) {
        const unsigned long long c = __ldg(cell_start + i + 1) - __ldg(cell_start + i);
        s += c * c;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

int pvi_build(pv_ctx *ctx, float maxdist, uint32_t nused) { return pvi_build_map(ctx, PV_MAP_VOLUME, maxdist, nused); }

// Build the lookup grid over one photon class: the volume photons (the context's main set; the medium's extent gate is folded
// into alpha) or one of the surface classes of pv_shoot_maps / pv_set_map_photons (no gate).  There is ONE grid per context.
int pvi_build_map(pv_ctx *ctx, int which, float maxdist, uint32_t nused) {
    ctx->built = false;
    if (which < PV_MAP_VOLUME || which > PV_MAP_RADIANCE) { ctx->err = "pv_select_map: bad map"; return PV_EINVAL; }
    const bool is_volume = which == PV_MAP_VOLUME;
    const float *src_pos = is_volume ? ctx->d_pos : ctx->surf[which - 1].pos, *src_wi = is_volume ? ctx->d_wi : ctx->surf[which - 1].wi;
    const float *src_alpha = is_volume ? ctx->d_alpha : ctx->surf[which - 1].alpha;
    uint64_t n = is_volume ? ctx->n_photons : ctx->surf[which - 1].n;
    ctx->map_which = which; ctx->map_n = n;
    if (n > 0xFFFFFFF0ull) { ctx->err = "too many photons for 32-bit indices"; return PV_EINVAL; }
    if (!(maxdist > 0.f)) { ctx->err = "pv_build: maxdist must be > 0"; return PV_EINVAL; }
    GridParams g{};
    if (n == 0) {
        g.origin[0] = g.origin[1] = g.origin[2] = 0.f; g.h = maxdist; g.inv_h = 1.f / maxdist; g.hx = g.h; g.inv_hx = g.inv_h; g.xshift = 0;
        g.dims[0] = g.dims[1] = g.dims[2] = 1; g.xbits = 0; g.yzbits = 0; g.table_size = 1; g.margin = 0.f;
        if (ctx->table_cap < 2) {
            if (ctx->cell_start) cudaFree(ctx->cell_start);
            PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->cell_start, 2 * sizeof(uint32_t))); ctx->table_cap = 2;
        }
        PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->cell_start, 0, 2 * sizeof(uint32_t), ctx->stream));
        ctx->grid = g; ctx->built = true;
        return PV_OK;
    }
    // 1. bounds
    int h_bounds[6];
    {
        int init[6];
        float pinf = INFINITY, ninf = -INFINITY;
        int ip, in_; memcpy(&ip, &pinf, 4); memcpy(&in_, &ninf, 4);
        for (int a = 0; a < 3; ++a) { init[a] = ip; init[3 + a] = in_ ^ 0x7fffffff; }
        int rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, 64); if (rc) return rc;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->io2, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
        int blocks = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
        bbox_kernel<<<blocks, 256, 0, ctx->stream>>>(src_pos, n, (int *)ctx->io2);
        ctx->launches += 1;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(h_bounds, ctx->io2, sizeof(h_bounds), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = ordered_float(h_bounds[a]); mx[a] = ordered_float(h_bounds[3 + a]); }
    for (int a = 0; a < 3; ++a)
        if (!std::isfinite(mn[a]) || !std::isfinite(mx[a])) { ctx->err = "pv_build: non-finite photon position"; return PV_EINVAL; }
    // 2. cell size: h = min(maxdist, radius expected to hold nused photons at the mean density)
    double ext[3], vol = 1.0, maxext = 0.0, maxabs = 0.0;
    for (int a = 0; a < 3; ++a) {
        ext[a] = std::max((double)mx[a] - (double)mn[a], 1e-6);
        vol *= ext[a]; maxext = std::max(maxext, ext[a]);
        maxabs = std::max(maxabs, std::max(std::fabs((double)mn[a]), std::fabs((double)mx[a])));
    }
    double rho = (double)n / vol;
    double hk = std::cbrt(3.0 * (double)std::max<uint32_t>(nused, 1) / (4.0 * M_PI * rho));
    // a hair above maxdist, so that the 3x3x3 block provably holds every photon within maxdist (see one_shell_r)
    const double margin_est = 1.01e-4 * (double)maxdist + 4e-6 * (maxabs + maxext);
    // k-nearest regime (the cell expected to hold nused photons is smaller than maxdist).  The cell-batched gather's k-nearest mode
    // (nused <= 64, pv_cellgather.cu) searches a trial sphere of ~1.1 hk inside the 5 x 5 rows around the query's cell and wants
    // cells of 0.75 hk (measured on config 2, k = 50, 1 M photons: 17.7 ms per frame; 1.0 hk 19.3 ms, 1.5 hk 21.4 ms); the
    // warp-per-lookup search that larger nused falls back to is indifferent between 0.5 and 1.0 hk (48-57 ms) and keeps hk.
    double knn_cell = nused <= 64 ? 0.75 : 1.0;
    if (const char *e = getenv("PV_KNN_CELL")) knn_cell = std::max(0.25, std::min(2.0, atof(e)));       // tuning knob
    double h = std::min((double)maxdist + 3.0 * margin_est, hk * knn_cell);
    const int max_dim = 256;                                  // key bits <= 24 -> cell table <= 64 MiB
    h = std::max(h, maxext / (max_dim - 1));
    h = std::max(h, 1e-6);
    // Crowded maps (photons along a spot beam, around a light): when the k-nearest search is the warp-per-lookup one (nused > 64)
    // and the cell a photon sits in holds many times nused photons, halve the cells (at most twice) -- every lookup near the crowd
    // scans whole cells.  Measured on projectScene/pinkfloyd.pbrt at its shipped size (5 M photons, nused 500, 0.8 degree spot):
    // gather of a 512 x 512 frame 1122 ms at the mean-density cell, 619 ms at half, 611 ms at a quarter of it; the batched
    // volume calls of the rays behind the prism 1.43 / 0.84 / 0.59 s.  Maps of even density (BASELINE configs 2, 3) never get here.
    const bool adaptive = nused > 64 && !getenv("PV_KNN_CELL");
    const auto t_build0 = std::chrono::steady_clock::now();
    int attempts = 0;
    const double h_min = std::max(maxext / (max_dim - 1), 1e-6);
    uint32_t *skeys = nullptr, *svals = nullptr;
    size_t nn = (size_t)n;
    for (int attempt = 0;; ++attempt) {
        g.h = (float)h; g.inv_h = 1.f / g.h;
        int maxd = 1;
        for (int a = 0; a < 3; ++a) {
            g.origin[a] = mn[a];
            g.dims[a] = std::min(max_dim, std::max(1, (int)std::floor(ext[a] / h) + 1));
            maxd = std::max(maxd, g.dims[a]);
        }
        g.yzbits = ceil_log2(std::max(g.dims[1], g.dims[2]));
        // finer cells along x (up to two halvings) while the key stays within 24 (26) bits and cells are not much
        // more numerous than photons
        g.xshift = 0;
        {
            const int coarse = g.dims[0];
            const int max_key_bits = n >= (1ull << 25) ? 26 : 24;      // cell table <= 64 MiB, 256 MiB for very large maps
            int xs_max = 2;
            if (const char *e = getenv("PV_XSHIFT_MAX")) xs_max = std::max(0, std::min(4, atoi(e)));      // tuning knob
            for (int xs = 1; xs <= xs_max; ++xs) {
                const int fine = coarse << xs;
                if (ceil_log2(fine) + 2 * g.yzbits > max_key_bits) break;
                if ((double)fine * g.dims[1] * g.dims[2] > 8.0 * (double)n) break;
                g.xshift = xs;
            }
        }
        g.hx = g.h / (float)(1 << g.xshift); g.inv_hx = 1.f / g.hx;       // exact: a power of two
        g.dims[0] = std::max(1, (int)std::floor(ext[0] / (double)g.hx) + 1);
        g.xbits = ceil_log2(g.dims[0]);
        g.table_size = (uint32_t)1 << (g.xbits + 2 * g.yzbits);
        g.margin = (float)(1e-4 * h + 4e-6 * (maxabs + maxext));
        g.one_shell_r = g.h - 2.f * g.margin;            // lookups with r <= this never need more than the 3x3x3 block
        int key_bits = std::max(1, g.xbits + 2 * g.yzbits);

        // 3. keys + sort
        size_t need = nn * 4 * sizeof(uint32_t) + 256;
        int rc = pv_ensure(ctx, &ctx->scratch, &ctx->scratch_bytes, need); if (rc) return rc;
        uint32_t *keys = (uint32_t *)ctx->scratch, *vals = keys + nn, *keys_tmp = vals + nn, *vals_tmp = keys_tmp + nn;
        keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src_pos, n, g, keys, vals);
        ctx->launches += 1;
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        rc = pvi_sort_pairs_u32(ctx, keys, vals, keys_tmp, vals_tmp, n, key_bits, &skeys, &svals); if (rc) return rc;

        // 5. cell table
        if (ctx->table_cap < g.table_size + 1) {
            if (ctx->cell_start) cudaFree(ctx->cell_start);
            ctx->cell_start = nullptr; ctx->table_cap = 0;
            PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->cell_start, ((size_t)g.table_size + 1) * sizeof(uint32_t)));
            ctx->table_cap = g.table_size + 1;
        }
        cell_start_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(skeys, n, g.table_size, ctx->cell_start);
        ctx->launches += 1;
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        attempts = attempt + 1;
        if (!adaptive || attempt >= 2 || 0.5 * h < h_min) break;
        // occupancy of the cell a photon sits in, against what a lookup wants
        rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, 64); if (rc) return rc;
        PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->io2, 0, sizeof(unsigned long long), ctx->stream));
        occupancy_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->cell_start, g.table_size, (unsigned long long *)ctx->io2);
        ctx->launches += 1;
        unsigned long long sumsq = 0;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(&sumsq, ctx->io2, sizeof(sumsq), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        if ((double)sumsq / (double)n <= 8.0 * (double)nused) break;
        h *= 0.5;
    }

    // 4. records in sorted order
    if (ctx->map_cap < n) {
        if (ctx->m_pos4) cudaFree(ctx->m_pos4);
        if (ctx->m_wi4) cudaFree(ctx->m_wi4);
        if (ctx->m_alpha32) cudaFree(ctx->m_alpha32);
        if (ctx->m_orig) cudaFree(ctx->m_orig);
        ctx->m_pos4 = nullptr; ctx->m_wi4 = nullptr; ctx->m_alpha32 = nullptr; ctx->m_orig = nullptr; ctx->map_cap = 0;
        PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->m_pos4, nn * sizeof(float4)));
        PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->m_wi4, nn * sizeof(float4)));
        PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->m_alpha32, nn * 32 * sizeof(float)));
        PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->m_orig, nn * sizeof(uint32_t)));
        ctx->map_cap = n;
    }
    {
        int blocks = (int)std::min<uint64_t>((n + 31) / 32, (uint64_t)ctx->sm_count * 16);
        gather_records_kernel<<<blocks, 256, 0, ctx->stream>>>(svals, n, src_pos, src_wi, src_alpha, ctx->m_pos4, ctx->m_wi4,
                                                             ctx->m_alpha32, ctx->m_orig, ctx->has_scene && is_volume ? ctx->dscene : nullptr);
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (getenv("PV_TIMING"))
        fprintf(stderr, "[pv timing] build of map %d: %llu photons, cell %.4g (x %.4g), %d x %d x %d cells, table 2^%d, %d pass(es), %.1f ms\n", which,
                (unsigned long long)n, (double)g.h, (double)g.hx, g.dims[0], g.dims[1], g.dims[2], g.xbits + 2 * g.yzbits, attempts,
                1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build0).count());
    ctx->grid = g;
    ctx->built = true;
    return PV_OK;
}
