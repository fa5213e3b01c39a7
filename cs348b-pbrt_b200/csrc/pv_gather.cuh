// pv_gather.cuh -- declarations shared by the gather kernels (pv_gather.cu: one warp per ray / per march step;
// pv_cellgather.cu: march steps sorted by photon-grid cell, one warp per batch of 32 neighbouring steps).
#pragma once
#include "pv_grid.cuh"
#include "pv_march.cuh"

struct MapView {
    const float4 *pos4;            // x, y, z, sorted position of the photon (bits): a staged candidate carries its own address
    const float4 *wi4; const float *alpha32; const uint32_t *cell_start;
    const uint32_t *orig;          // sorted position -> original photon index (tie-breaks and the k-NN output only)
    GridParams g;
    uint64_t n;
    int need_wi;                   // the medium's phase function depends on wi (g != 0)
};

// ---- mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(mbar), "r"(phase) : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

struct GatherArgs {
    MapView m;
    const DevScene *sc;
    const pv_ray *rays;
    const RayHdr *hdr;             // per ray: where its march steps are (pv_march.cu)
    const StepRec *steps;
    uint64_t n;
    float maxdist;
    uint32_t nused, flags, cap;
    float *L, *T;
    pv_gather_stats *stats;
    unsigned long long *counter;
    // step-parallel ("latency") form: in-scattered radiance of every march step, 32 floats per StepRec, filled by
    // gather_lii_kernel and consumed by gather_kernel<true>
    float *lii; unsigned long long total_steps;
    // gather_lii_kernel over a LIST of steps (the cell-batched gather's overflow): list[0 .. *list_count)
    const uint32_t *list; const unsigned long long *list_count;
    const uint32_t *block_order;   // recurrence_thread_kernel: CTA i takes the rays of block block_order[i] (pv_march.cu block_order_kernel), or null
};


MapView pvi_map_view(pv_ctx *ctx);
// pv_cellgather.cu: fills a.lii (32 floats per march step) for every live step of the slice whose neighbour count stays within
// a.nused; the steps it could not finish (more than nused photons in range: the k-nearest selection is needed) are appended to
// ctx->cg_overflow (count in ctx->d_counters[CG_CNT_OVERFLOW]) for gather_lii_kernel.
enum { CG_CNT_BATCH = 16, CG_CNT_OVERFLOW = 17, CG_CNT_BATCH2 = 18, CG_CNT_OVERFLOW2 = 19 };
int pvi_cellgather(pv_ctx *ctx, const GatherArgs &a, const uint32_t **leftover_list, int *leftover_count);
