// pv_lbvh.cu -- the scene BVH built on the device ("GPU LBVH build replacing CPU SAH/AAC").
//
// Replaces the BVHAccel constructor of the reference (accelerators/bvh.cpp:196-300: primitive bounds -> recursiveBuild :301-556
// -> flattenBVHTree :559-577) for scenes whose aggregate is not a BVHAccel, or whose mesh is large enough that the reference's
// serial build is what the user waits for.  The OUTPUT is the reference's own structure: an array of 32-byte LinearBVHNode records
// in depth-first order (first child = node + 1, second child at secondChildOffset, interior nodes carry the axis along which the
// first child lies on the low side, leaves carry a range of the reordered primitive array), so bvh_traverse (pv_device.cuh, the
// restatement of bvh.cpp:585-685) walks it unchanged and pv_set_scene validates it like an exported one.
//
// The build is a linear BVH over Morton codes (Lauterbach et al. 2009) with the hierarchy of Karras 2012 ("Maximizing parallelism
// in the construction of BVHs, octrees and k-d trees"):
//   lbvh_centroid_bounds_kernel  bound of the primitive centroids                  (bvh.cpp:322-325 computes the same per node)
//   lbvh_morton_kernel           30-bit Morton code of every centroid in that bound
//   rs_count / rs_scatter        the map build's LSD radix sort (pv_build.cu), stable: ties keep primitive order
//   lbvh_hierarchy_kernel        one thread per interior node: its key range and split from common-prefix lengths
//   lbvh_refit_kernel            one thread per leaf climbs to the root; the second arrival at a node unions the child bounds and
//                                counts the nodes its subtree will EMIT (a subtree of <= max_prims_in_node primitives is one leaf,
//                                bvh.cpp:334-346 makes leaves of small ranges too)
//   lbvh_emit_kernel             one thread per tree node: depth-first index = sum over its ancestors of (1, or 1 + the emitted
//                                size of the left sibling's subtree) -- no serial flattening pass
// All of it is O(n) memory traffic plus a 4-pass sort of 8-byte pairs: HBM-bound, ~32 n + 4 * 16 n + 2 * 64 n bytes.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "pv_ctx.h"

#define LB_THREADS 256
#define LB_LEAF 0x80000000u

__device__ __forceinline__ int lb_float_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
static inline float lb_ordered_float(int i) { int j = i >= 0 ? i : i ^ 0x7fffffff; float f; memcpy(&f, &j, 4); return f; }

__device__ __forceinline__ float lb_centroid(const float *b, int a) { return 0.5f * b[a] + 0.5f * b[3 + a]; }   // bvh.cpp:51

__global__ void __launch_bounds__(LB_THREADS) lbvh_centroid_bounds_kernel(const float *__restrict__ pb, uint32_t n, int *__restrict__ out) {
    __shared__ int smin[3][LB_THREADS / 32], smax[3][LB_THREADS / 32];
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float *b = pb + 6 * (size_t)i;
#pragma unroll
        for (int a = 0; a < 3; ++a) { int c = lb_float_ordered(lb_centroid(b, a)); lo[a] = min(lo[a], c); hi[a] = max(hi[a], c); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) { lo[a] = min(lo[a], __shfl_xor_sync(PV_FULL, lo[a], o)); hi[a] = max(hi[a], __shfl_xor_sync(PV_FULL, hi[a], o)); }
        if ((threadIdx.x & 31) == 0) { smin[a][threadIdx.x >> 5] = lo[a]; smax[a][threadIdx.x >> 5] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int a = threadIdx.x, l = smin[a][0], h = smax[a][0];
        for (int w = 1; w < LB_THREADS / 32; ++w) { l = min(l, smin[a][w]); h = max(h, smax[a][w]); }
        atomicMin(out + a, l); atomicMax(out + 3 + a, h);
    }
}

__device__ __forceinline__ uint32_t lb_expand10(uint32_t v) {        // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ float lb_ordered_float_dev(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__global__ void __launch_bounds__(LB_THREADS) lbvh_morton_kernel(const float *__restrict__ pb, uint32_t n, const int *__restrict__ cb,
                                                                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *b = pb + 6 * (size_t)i;
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = lb_ordered_float_dev(__ldg(cb + a)), hi = lb_ordered_float_dev(__ldg(cb + 3 + a));
        float f = hi > lo ? (lb_centroid(b, a) - lo) * __fdiv_rn(1024.f, hi - lo) : 0.f;      // [0, 1024]
        q[a] = (uint32_t)min(max(f, 0.f), 1023.f);
    }
    keys[i] = (lb_expand10(q[0]) << 2) | (lb_expand10(q[1]) << 1) | lb_expand10(q[2]);     // bit 29 = x's top bit, 28 = y's, 27 = z's
    vals[i] = i;
}

// Karras 2012, section 4: delta(i, j) = length of the common prefix of keys i and j, -1 outside the array; equal keys are told apart
// by their positions (the sort is stable, so that is primitive order)
__device__ __forceinline__ int lb_delta(const uint32_t *__restrict__ keys, int n, int i, uint32_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t kj = __ldg(keys + j);
    return ki == kj ? 32 + __clz((uint32_t)i ^ (uint32_t)j) : __clz(ki ^ kj);
}
// Tree node ids: interior i in [0, n-1), leaf j is (n - 1) + j.  Child links carry LB_LEAF for leaves (then the low bits are j).
__global__ void __launch_bounds__(LB_THREADS) lbvh_hierarchy_kernel(const uint32_t *__restrict__ keys, int n, uint32_t *__restrict__ left,
                                                                     uint32_t *__restrict__ right, uint32_t *__restrict__ parent,
                                                                     uint2 *__restrict__ range) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    uint32_t ki = __ldg(keys + i);
    int d = lb_delta(keys, n, i, ki, i + 1) - lb_delta(keys, n, i, ki, i - 1) >= 0 ? 1 : -1;
    int dmin = lb_delta(keys, n, i, ki, i - d);
    int lmax = 2;
    while (lb_delta(keys, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lb_delta(keys, n, i, ki, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lb_delta(keys, n, i, ki, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lb_delta(keys, n, i, ki, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    uint32_t lc = lo == gamma ? (LB_LEAF | (uint32_t)gamma) : (uint32_t)gamma;
    uint32_t rc = hi == gamma + 1 ? (LB_LEAF | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    left[i] = lc; right[i] = rc;
    range[i] = make_uint2((uint32_t)lo, (uint32_t)hi);
    parent[(lc & LB_LEAF) ? (n - 1) + (lc & ~LB_LEAF) : lc] = (uint32_t)i;
    parent[(rc & LB_LEAF) ? (n - 1) + (rc & ~LB_LEAF) : rc] = (uint32_t)i;
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

struct LbBox { float4 a; float2 b; };            // 24 bytes: min.xyz max.x | max.yz
__device__ __forceinline__ uint32_t lb_node_id(uint32_t link, int n) { return (link & LB_LEAF) ? (uint32_t)(n - 1) + (link & ~LB_LEAF) : link; }

__global__ void __launch_bounds__(LB_THREADS) lbvh_refit_kernel(const float *__restrict__ pb, const uint32_t *__restrict__ order, int n,
                                                                 const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
                                                                 const uint32_t *__restrict__ parent, const uint2 *__restrict__ range,
                                                                 uint32_t max_prims, float *box /* 6 per tree node */, uint32_t *esize,
                                                                 uint32_t *flag) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t node = (uint32_t)(n - 1) + (uint32_t)j;
    const float *b = pb + 6 * (size_t)__ldg(order + j);
    float bb[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { bb[k] = __ldg(b + k); __stcg(box + 6 * (size_t)node + k, bb[k]); }
    __stcg(esize + node, 1u);
    uint32_t cur = __ldg(parent + node);
    while (cur != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(flag + cur, 1u) == 0u) return;          // the sibling subtree is not done: its thread will pass here later
        __threadfence();                                         // acquire side: the sibling's box / size stores precede its atomic
        uint32_t l = lb_node_id(__ldg(left + cur), n), r = lb_node_id(__ldg(right + cur), n);
#pragma unroll
        for (int k = 0; k < 3; ++k) {                            // Union(Bounds, Bounds) core/geometry.cpp:53-63
            bb[k] = fminf(__ldcg(box + 6 * (size_t)l + k), __ldcg(box + 6 * (size_t)r + k));
            bb[3 + k] = fmaxf(__ldcg(box + 6 * (size_t)l + 3 + k), __ldcg(box + 6 * (size_t)r + 3 + k));
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) __stcg(box + 6 * (size_t)cur + k, bb[k]);
        uint2 rg = __ldg(range + cur);
        uint32_t cnt = rg.y - rg.x + 1u;
        __stcg(esize + cur, cnt <= max_prims ? 1u : 1u + __ldcg(esize + l) + __ldcg(esize + r));
        cur = __ldg(parent + cur);
    }
}

__global__ void __launch_bounds__(LB_THREADS) lbvh_emit_kernel(const uint32_t *__restrict__ keys, int n, const uint32_t *__restrict__ left,
                                                                const uint32_t *__restrict__ right, const uint32_t *__restrict__ parent,
                                                                const uint2 *__restrict__ range, const float *__restrict__ box,
                                                                const uint32_t *__restrict__ esize, uint32_t max_prims,
                                                                pv_bvh_node *__restrict__ out) {
    uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= (uint32_t)(2 * n - 1)) return;
    const bool is_leaf = node >= (uint32_t)(n - 1);
    uint32_t first, count;
    if (is_leaf) { first = node - (uint32_t)(n - 1); count = 1; }
    else { uint2 rg = __ldg(range + node); first = rg.x; count = rg.y - rg.x + 1u; }
    uint32_t p = __ldg(parent + node);
    if (p != 0xFFFFFFFFu) {                                       // inside a subtree that became one leaf: not emitted
        uint2 prg = __ldg(range + p);
        if (prg.y - prg.x + 1u <= max_prims) return;
    }
    uint32_t idx = 0;
    for (uint32_t cur = node; p != 0xFFFFFFFFu; cur = p, p = __ldg(parent + p)) {
        uint32_t l = lb_node_id(__ldg(left + p), n);
        idx += (l == cur) ? 1u : 1u + __ldg(esize + l);
    }
    float4 a; float4 b;
    const float *bx = box + 6 * (size_t)node;
    a = make_float4(bx[0], bx[1], bx[2], bx[3]);
    b.x = bx[4]; b.y = bx[5];
    if (count <= max_prims) {                                     // leaf: primitivesOffset, nPrimitives (bvh.cpp:566-569)
        b.z = __uint_as_float(first);
        b.w = __uint_as_float(count & 0xffu);
    } else {                                                      // interior: secondChildOffset, axis (bvh.cpp:570-575)
        uint32_t l = lb_node_id(__ldg(left + node), n);
        uint32_t split = first + ((__ldg(left + node) & LB_LEAF) ? 0u : (__ldg(range + l).y - first));   // last key of the left child
        uint32_t k0 = __ldg(keys + split), k1 = __ldg(keys + split + 1);
        uint32_t axis = k0 == k1 ? 0u : (uint32_t)(__clz(k0 ^ k1) - 2) % 3u;
        b.z = __uint_as_float(idx + 1u + __ldg(esize + l));
        b.w = __uint_as_float(axis << 8);
    }
    float4 *o = reinterpret_cast<float4 *>(out + idx);
    o[0] = a; o[1] = b;
}

int pvi_build_bvh(pv_ctx *ctx, const float *prim_bounds, uint32_t n, uint32_t max_prims, pv_bvh_node *nodes, uint32_t nodes_cap,
                  uint32_t *n_nodes, uint32_t *prim_order, float *device_ms) {
    *n_nodes = 0;
    if (device_ms) *device_ms = 0.f;
    if (n == 0) return PV_OK;
    if (n > 0x10000000u) { ctx->err = "pv_build_bvh: more than 2^28 primitives"; return PV_EINVAL; }
    if (n == 1) {                                                 // a tree of one leaf (bvh.cpp:334-346 with nPrimitives == 1)
        if (nodes_cap < 1) { ctx->err = "pv_build_bvh: nodes_cap too small"; return PV_EINVAL; }
        memset(&nodes[0], 0, sizeof(pv_bvh_node));
        memcpy(nodes[0].bounds, prim_bounds, 6 * sizeof(float));
        nodes[0].offset = 0; nodes[0].n_primitives = 1;
        prim_order[0] = 0; *n_nodes = 1;
        return PV_OK;
    }
    const size_t nt = 2 * (size_t)n - 1;                          // tree nodes before small subtrees collapse
    // one allocation for the whole build (a scene is built once): bounds in, keys / values x2, links, boxes, sizes, flags, nodes out
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_pb = take(6 * sizeof(float) * n), o_k0 = take(4 * (size_t)n), o_k1 = take(4 * (size_t)n), o_v0 = take(4 * (size_t)n),
                 o_v1 = take(4 * (size_t)n), o_left = take(4 * (size_t)n), o_right = take(4 * (size_t)n), o_parent = take(4 * nt),
                 o_range = take(8 * (size_t)n), o_box = take(24 * nt), o_esize = take(4 * nt), o_flag = take(4 * (size_t)n),
                 o_cb = take(32), o_out = take(sizeof(pv_bvh_node) * nt);
    char *base = nullptr;
    if (cudaMalloc(&base, off) != cudaSuccess) { cudaGetLastError(); ctx->err = "pv_build_bvh: out of device memory"; return PV_ENOMEM; }
    struct Free { char *p; ~Free() { cudaFree(p); } } guard{base};
    float *d_pb = (float *)(base + o_pb);
    uint32_t *k0 = (uint32_t *)(base + o_k0), *k1 = (uint32_t *)(base + o_k1), *v0 = (uint32_t *)(base + o_v0), *v1 = (uint32_t *)(base + o_v1);
    uint32_t *left = (uint32_t *)(base + o_left), *right = (uint32_t *)(base + o_right), *parent = (uint32_t *)(base + o_parent);
    uint2 *range = (uint2 *)(base + o_range);
    float *box = (float *)(base + o_box);
    uint32_t *esize = (uint32_t *)(base + o_esize), *flag = (uint32_t *)(base + o_flag);
    int *d_cb = (int *)(base + o_cb);
    pv_bvh_node *d_out = (pv_bvh_node *)(base + o_out);
    cudaStream_t st = ctx->stream;

    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pb, prim_bounds, 6 * sizeof(float) * n, cudaMemcpyHostToDevice, st));
    const int cb_init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cb, cb_init, sizeof(cb_init), cudaMemcpyHostToDevice, st));
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(flag, 0, 4 * (size_t)n, st));
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev2, st));
    const uint32_t gb = (n + LB_THREADS - 1) / LB_THREADS;
    lbvh_centroid_bounds_kernel<<<std::min<uint32_t>(gb, 4u * (uint32_t)ctx->sm_count), LB_THREADS, 0, st>>>(d_pb, n, d_cb);
    lbvh_morton_kernel<<<gb, LB_THREADS, 0, st>>>(d_pb, n, d_cb, k0, v0);
    uint32_t *sk, *sv;
    int rc = pvi_sort_pairs_u32(ctx, k0, v0, k1, v1, n, 30, &sk, &sv); if (rc) return rc;
    lbvh_hierarchy_kernel<<<(n - 1 + LB_THREADS - 1) / LB_THREADS, LB_THREADS, 0, st>>>(sk, (int)n, left, right, parent, range);
    lbvh_refit_kernel<<<gb, LB_THREADS, 0, st>>>(d_pb, sv, (int)n, left, right, parent, range, max_prims, box, esize, flag);
    lbvh_emit_kernel<<<(uint32_t)((nt + LB_THREADS - 1) / LB_THREADS), LB_THREADS, 0, st>>>(sk, (int)n, left, right, parent, range, box, esize,
                                                                                                max_prims, d_out);
    ctx->launches += 5;
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev3, st));
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    uint32_t total = 0;
    int cbi[6];
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(&total, esize, 4, cudaMemcpyDeviceToHost, st));      // emitted size of the root's subtree
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(cbi, d_cb, sizeof(cbi), cudaMemcpyDeviceToHost, st));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    for (int a = 0; a < 6; ++a) {
        float f = lb_ordered_float(cbi[a]);
        if (!(f == f) || std::isinf(f)) { ctx->err = "pv_build_bvh: a primitive bound is not finite"; return PV_EINVAL; }
    }
    if (total == 0 || total > nt) { ctx->err = "pv_build_bvh: internal error (node count)"; return PV_ECUDA; }
    if (total > nodes_cap) { ctx->err = "pv_build_bvh: nodes_cap too small (2 * n_prims - 1 always suffices)"; return PV_EINVAL; }
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(nodes, d_out, sizeof(pv_bvh_node) * (size_t)total, cudaMemcpyDeviceToHost, st));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(prim_order, sv, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    if (device_ms) PV_CUDA_CHECK(ctx, cudaEventElapsedTime(device_ms, ctx->ev2, ctx->ev3));
    *n_nodes = total;
    return PV_OK;
}
