// pv_ctx.h -- host-side context behind the C ABI (include/pv.h).
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <string>
#include <vector>
#include "pv_device.cuh"

// Grid of photon cells.  Sort key of a cell = (morton2(cy, cz) << xbits) | cx:
// rows of cells along x are contiguous in memory (the part of a row a lookup needs is ONE contiguous
// photon range), rows themselves follow a Morton (Z-order) curve over (y, z).  Cells are h x h in (y, z) and
// hx = h / 2^xshift along x: the finer x resolution lets a lookup clip each row to the chord of its search
// sphere at no extra cost (still one range per row).  "Coarse" x cells (2^xshift fine cells, h wide) keep the
// shell geometry of the k-nearest search cubic.
struct GridParams {
    float origin[3];
    float h, inv_h;
    float hx, inv_hx;       // cell size along x
    int xshift;             // h = hx * 2^xshift
    int dims[3];            // dims[0] counts FINE x cells
    int xbits, yzbits;      // key bits: xbits + 2*yzbits
    uint32_t table_size;    // number of keys (cell_start has table_size + 1 entries)
    float margin;           // conservative slack subtracted from the guaranteed search radius
    float one_shell_r;      // radius up to which the 3x3x3 block of an in-grid query is exhaustive
};

// a photon set in deposit order (same SoA planes as the context's main set)
struct PhotonSet {
    float *pos = nullptr, *wi = nullptr, *alpha = nullptr;   // pos[3n], wi[3n], alpha[32n]
    uint64_t *ids = nullptr;
    uint64_t n = 0, cap = 0;
};

struct pv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // pv_gather overlaps its host copies with the kernels
    std::mutex mu;
    std::string err;

    // scene
    DevScene hscene{};
    DevScene *dscene = nullptr;
    void *d_nodes = nullptr, *d_tri = nullptr, *d_prim_mat = nullptr, *d_mats = nullptr, *d_lights = nullptr, *d_density = nullptr;
    void *d_spheres = nullptr, *d_mat_flags = nullptr, *d_ltris = nullptr, *d_ltri_area = nullptr, *d_ltri_cdf = nullptr;
    bool has_scene = false;

    // photons in deposit order, SoA planes: pos[3n], wi[3n], alpha[32n] (30 bins + 2 pad = one 128-byte line,
    // so the shooter stores and the map build moves it with 128-bit accesses), ids[n]
    float *d_pos = nullptr, *d_wi = nullptr, *d_alpha = nullptr;
    uint64_t *d_ids = nullptr;
    uint64_t n_photons = 0, cap_photons = 0;

    // surface maps of the shooter (core/photonshooter.cpp:147-189): caustic, indirect, direct photons and radiance-photon
    // sites (wi plane = faceforwarded normal, alpha plane = rho_r), indexed by PV_MAP_* - 1; Lo of the radiance photons
    PhotonSet surf[4];
    float *rad_Lo = nullptr; uint64_t rad_Lo_cap = 0; bool rad_valid = false;
    uint64_t map_paths[4] = {0, 0, 0, 0};          // nCausticPaths, nIndirectPaths, nDirectPaths, nVolumePaths
    int map_which = 0; uint64_t map_n = 0;         // photon class (PV_MAP_*) and size of the set the grid was last built over

    // the map: photons sorted by cell key
    float4 *m_pos4 = nullptr;      // x, y, z, sorted position (bits)
    uint32_t *m_orig = nullptr;    // sorted position -> original photon index
    float4 *m_wi4 = nullptr;       // wi.xyz, 0
    float *m_alpha32 = nullptr;    // 32 floats per photon: alpha[30], 0, 0  (one 128-byte line)
    uint32_t *cell_start = nullptr;
    uint64_t map_cap = 0; uint32_t table_cap = 0;
    GridParams grid{};
    bool built = false;

    // scratch for sort / staging
    void *scratch = nullptr; size_t scratch_bytes = 0;
    void *sort_hist = nullptr; size_t sort_hist_bytes = 0;   // digit histograms of the radix sort
    void *io = nullptr; size_t io_bytes = 0;       // device staging for host-pointer entry points
    void *io2 = nullptr; size_t io2_bytes = 0;

    double shoot_yield[2] = {0., 0.};              // deposits per light path seen so far (volume-only pass, all-maps pass): sizes the next wave's buffer
    // pv_gather_indexed / pv_volume_li_indexed: stream index of every ray of the call in flight (device), and the ray array it belongs to
    const uint64_t *d_ray_index = nullptr; const pv_ray *ray_index_rays = nullptr;
    void *io3 = nullptr; size_t io3_bytes = 0;
    uint64_t wf_deep_pages = 0;                    // size of the wavefront's deep-stack page pool (grown when a wave runs it dry)
    void *wf = nullptr; size_t wf_bytes = 0;       // slot state of the shooter's wavefront (pv_wavefront.cu)

    pv_gather_stats *d_stats = nullptr;
    unsigned long long *d_counters = nullptr;      // work-distribution counters
    // march records of the ray slice being gathered (pv_march.cu): RayHdr per ray, StepRec per march step
    void *march_hdr = nullptr; size_t march_hdr_bytes = 0;
    void *march_steps = nullptr; size_t march_steps_bytes = 0;
    void *march_blk = nullptr; size_t march_blk_bytes = 0;      // per block of 128 rays of the slice: its march steps, and the blocks in order of decreasing steps
    void *lii = nullptr; size_t lii_bytes = 0;     // per-step in-scattered radiance of the step-parallel gather (32 floats per step)
    void *cg_sort = nullptr; size_t cg_sort_bytes = 0;           // cell-batched gather: (cell key, step) pairs and their sort buffers
    void *cg_overflow = nullptr; size_t cg_overflow_bytes = 0;   // steps left to the warp-per-step kernel (more than nused photons in range)
    cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};   // after sort / after cellgather_kernel / after the overflow pass
    float phase_ms[4] = {0.f, 0.f, 0.f, 0.f};                    // last pv_gather: step sort, cellgather_kernel, overflow pass, recurrence
    unsigned long long *h_total = nullptr;         // mapped pinned word: step count of the slice
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // around gather_kernel
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;      // around the march kernels
    float last_ms = 0.f, last_march_ms = 0.f;
    int sm_count = 148;
    void *comm = nullptr; int comm_rank = 0, comm_world = 0;     // ncclComm_t of pv_comm_init / pv_comm_init_all (pv_comm.cu)
    uint64_t launches = 0;                         // kernels launched on the build / gather path (pv_launch_count)
};

#define PV_CUDA_CHECK(ctx, call)                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
            return PV_ECUDA;                                                                       \
        }                                                                                          \
    } while (0)

int pv_ensure(pv_ctx *ctx, void **p, size_t *cap, size_t bytes);

// pv_build.cu
int pvi_build(pv_ctx *ctx, float maxdist, uint32_t nused);
int pvi_build_map(pv_ctx *ctx, int which, float maxdist, uint32_t nused);
int pvi_sort_pairs_u32(pv_ctx *ctx, uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int key_bits,
                       uint32_t **keys_out, uint32_t **vals_out);
int pvi_sort_pairs_u64(pv_ctx *ctx, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int key_bits,
                       uint64_t **keys_out, uint32_t **vals_out);
// pv_march.cu
#define PV_MARCH_MAX_BYTES (16ull << 30)          // step records of one ray slice
#define PV_GATHER_SLICE_RAYS (4ull << 20)
#define PV_GATHER_MAX_SLICES 8                     // host-pointer pv_gather: copy/compute pipeline depth
int pvi_march(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, uint32_t flags, uint64_t *total_steps);
// pv_gather.cu
int pvi_knn(pv_ctx *ctx, const float *d_pts, uint64_t n, uint32_t k, float r2, uint32_t *d_idx, float *d_d2, uint32_t *d_nfound);
int pvi_lphoton(pv_ctx *ctx, const float *d_pts, const float *d_w, uint64_t n, uint32_t nused, float maxdist, float *d_L);
int pvi_gather(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, float *d_L, float *d_T);
// pv_volint.cu
int pvi_volume_li(pv_ctx *ctx, int integrator, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, float *d_L, float *d_T);
// pv_trace.cu
int pvi_intersect(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, uint32_t *d_prim, float *d_t);
int pvi_occluded(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, uint8_t *d_hit);
int pvi_transmittance(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, float step, const float *d_u, float *d_T);
// pv_lbvh.cu
int pvi_build_bvh(pv_ctx *ctx, const float *prim_bounds, uint32_t n, uint32_t max_prims, pv_bvh_node *nodes, uint32_t nodes_cap,
                  uint32_t *n_nodes, uint32_t *prim_order, float *device_ms);
// pv_api.cu
int pvi_reserve_photons(pv_ctx *ctx, uint64_t n);
// pv_comm.cu
int pvi_comm_unique_id(uint8_t *id, std::string *err);
int pvi_comm_init(pv_ctx *ctx, const uint8_t *id, int rank, int world);
int pvi_comm_init_all(pv_ctx **ctxs, int n);
int pvi_comm_destroy(pv_ctx *ctx);
int pvi_allgather_photons(pv_ctx *ctx, int renumber, float *collective_ms);
int pvi_broadcast_photons(pv_ctx **ctxs, int n, int src, float *collective_ms);
// pv_shoot.cu
int pvi_shoot_blocks(pv_ctx *ctx, uint64_t first_block, uint32_t n_blocks, const pv_shoot_params *prm, uint32_t *counts, pv_shoot_stats *stats);
int pvi_shoot_finish(pv_ctx *ctx, uint64_t last_block);
int pvi_shoot(pv_ctx *ctx, uint64_t n_wanted, const pv_shoot_params *prm, pv_shoot_stats *stats);
int pvi_shoot_maps(pv_ctx *ctx, const pv_maps_params *mp, const pv_shoot_params *prm, pv_allreduce_u32_fn allreduce, void *user,
                   pv_maps_stats *out);
int pvi_reserve_set(pv_ctx *ctx, PhotonSet *s, uint64_t n);
void pvi_free_set(PhotonSet *s);
// pv_gather.cu
int pvi_radiance(pv_ctx *ctx, uint32_t n_lookup, float max_dist2, const uint64_t counts[3]);
int pvi_surface_lphoton(pv_ctx *ctx, const float *d_pts, const float *d_nf, uint64_t n, uint32_t n_lookup, float max_dist2, uint64_t n_paths,
                        float *d_Lr, float *d_Lt);
int pvi_radiance_nearest(pv_ctx *ctx, const float *d_pts, const float *d_n, uint64_t n, uint32_t *d_idx, float *d_Lo30);
int pvi_final_gather(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, float step, uint64_t seed, uint64_t index_base, float *d_Lindir, uint32_t *d_idx);
