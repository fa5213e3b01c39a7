/* pv.h -- C ABI of the B200-native volumetric photon-mapping path.
 *
 * This is the drop-in boundary for the hot path of piwell/CS348B-pbrt (a pbrt-v2
 * fork): photon shooting (core/photonshooter.cpp:47-357), the photon map
 * (core/kdtree.h:99-183) and the per-camera-ray gather
 * (integrators/photonvolume.cpp:15-222).  The reference has no FFI; the host
 * adapter classes that keep the PhotonShooter / PhotonVolumeIntegrator C++
 * signatures (core/photonshooter.h:81-116, integrators/photonvolume.h:14-38)
 * call exactly these entry points.
 *
 * Conventions
 *  - plain C types, host pointers unless the name ends in _dev (then the
 *    pointers are CUDA device addresses on the context's device);
 *  - every call returns 0 on success or a negative PV_E* code; the message is
 *    available from pv_last_error(); nothing throws;
 *  - a pv_ctx owns all device memory; calls on one ctx are serialised by an
 *    internal mutex (the reference calls Li/Transmittance from N pthreads,
 *    core/parallel.cpp:728-737);
 *  - there is NO CPU fallback: without a CUDA device pv_create fails.
 *  - spectra are the reference's 30-bin SampledSpectrum coefficients
 *    (core/spectrum.h:44-46), WITHOUT the fork's lambda/intensity members.
 */
#ifndef PV_H
#define PV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_NSPEC 30

#define PV_OK          0
#define PV_EINVAL     -1   /* bad argument / missing scene or map            */
#define PV_ECUDA      -2   /* CUDA runtime error (message has the detail)    */
#define PV_ENOMEM     -3
#define PV_ESTATE     -4   /* call order: e.g. gather before build           */
#define PV_ENOPHOTONS -5   /* reference's "Unable to store enough photons"   */

typedef struct pv_ctx pv_ctx;

/* accelerators/bvh.cpp:154-164 LinearBVHNode, same 32-byte layout. */
typedef struct pv_bvh_node {
    float    bounds[6];      /* pMin.xyz, pMax.xyz                           */
    uint32_t offset;         /* leaf: primitivesOffset; interior: secondChildOffset */
    uint8_t  n_primitives;   /* 0 -> interior                                */
    uint8_t  axis;
    uint8_t  pad[2];
} pv_bvh_node;

/* core/geometry.h:322-345 Ray (+ the scatter jitter sample->oneD[..][0] that
 * integrators/photonvolume.cpp:135 reads). 40 bytes. */
typedef struct pv_ray {
    float o[3];
    float d[3];
    float mint, maxt;
    float time;
    float u_scatter;
} pv_ray;

/* PV_LIGHT_AREA: DiffuseAreaLight over a triangle mesh (lights/diffuse.cpp:39-106, its ShapeSet core/light.cpp:114-172):
 * intensity[] = Lemit, `area` names its triangles in pv_scene_desc::light_tris (the ShapeSet's own refine order -- the order the
 * area distribution samples by).  The same triangles are ordinary primitives of the scene's BVH.                       */
enum { PV_LIGHT_POINT = 0, PV_LIGHT_SPOT = 1, PV_LIGHT_DISTANT = 2, PV_LIGHT_AREA = 3 };
#define PV_AREA_REVERSE_ORIENTATION 1u
#define PV_AREA_SWAPS_HANDEDNESS 2u
typedef struct pv_light {
    int32_t type;
    union {
        float pos[3];               /* point/spot lightPos (lights/point.cpp:43) */
        struct { uint32_t first_tri, n_tris, flags; } area;     /* PV_LIGHT_AREA: range in light_tris, PV_AREA_* flags */
    };
    float   dir[3];                 /* distant lightDir (lights/distant.cpp:43)  */
    float   cos_total_width;        /* spot (lights/spot.cpp:45-46)              */
    float   cos_falloff_start;
    float   intensity[PV_NSPEC];    /* I (point/spot) or L (distant)             */
    float   light_to_world[16];     /* row-major Matrix4x4, spot only            */
    float   world_to_light[16];
    float   power_y;                /* Light::Power(scene).y(), for the CDF
                                       (core/integrator.cpp:261-268)             */
} pv_light;

enum { PV_MEDIUM_NONE = 0, PV_MEDIUM_HOMOGENEOUS = 1, PV_MEDIUM_GRID = 2,
       PV_MEDIUM_RAINBOW = 3, PV_MEDIUM_EXPONENTIAL = 4 };
typedef struct pv_medium {
    int32_t type;
    float   world_to_volume[16];    /* row-major                                  */
    float   p0[3], p1[3];           /* extent (volumes/homogeneous.h:87)          */
    float   sigma_a[PV_NSPEC], sigma_s[PV_NSPEC], le[PV_NSPEC];
    float   g;
    int32_t nx, ny, nz;             /* grid only (volumes/volumegrid.h:66-70)     */
    const float *density;           /* host, nx*ny*nz, index z*nx*ny + y*nx + x.
                                       PV_MEDIUM_EXPONENTIAL (volumes/exponential.h:43-68):
                                       nx = 5, ny = nz = 1, density = {a, b, updir.xyz}
                                       (updir normalised as the ctor leaves it)       */
} pv_medium;

enum { PV_MAT_MATTE = 0, PV_MAT_GLASS = 1 };
typedef struct pv_material {
    int32_t type;
    float   kd[PV_NSPEC];           /* matte Kd (materials/matte.cpp:42-63), sigma==0 only */
    float   kr[PV_NSPEC], kt[PV_NSPEC];
    float   index, vn;              /* glass (materials/glass.cpp:42-70)          */
} pv_material;

/* shapes/sphere.cpp:40-164 Sphere: a BVH primitive that is not a triangle.
 * Matrices row-major like pv_light's. */
typedef struct pv_sphere {
    float   object_to_world[16], world_to_object[16];
    float   radius, zmin, zmax;             /* clamped as the ctor leaves them (:44-46)  */
    float   theta_min, theta_max, phi_max;  /* radians (:47-49)                         */
    int32_t flip_normal;                    /* ReverseOrientation ^ TransformSwapsHandedness
                                               (core/diffgeom.cpp:52-53)                 */
} pv_sphere;
#define PV_SHAPE_TRIANGLE 0xFFFFFFFFu

/* Flattened scene, exported from the unchanged host objects. */
typedef struct pv_scene_desc {
    const pv_bvh_node *nodes;       uint32_t n_nodes;
    const float       *tri_verts;   /* 9 floats per primitive, world space, in the
                                       BVH's reordered primitive order            */
    const uint32_t    *prim_material;
    uint32_t           n_prims;
    const pv_material *materials;   uint32_t n_materials;
    const pv_light    *lights;      uint32_t n_lights;
    const pv_medium   *medium;      /* NULL -> no volume region                   */
    float              world_bound[6];
    float              cie_y[PV_NSPEC];  /* SampledSpectrum::Y bin averages
                                            (core/spectrum.h:368-381)             */
    /* optional: primitives that are spheres.  prim_shape[i] = PV_SHAPE_TRIANGLE or an index
     * into spheres[]; NULL / n_spheres == 0 -> every primitive is a triangle (the
     * tri_verts slot of a sphere primitive is ignored).                           */
    const uint32_t    *prim_shape;
    const pv_sphere   *spheres;     uint32_t n_spheres;
    /* optional: the triangles of the area lights, 9 floats each, world space (see PV_LIGHT_AREA) */
    const float       *light_tris;  uint32_t n_light_tris;
} pv_scene_desc;

/* PhotonVolumeIntegrator ctor params (integrators/photonvolume.h:17-20) +
 * the counter-based RNG key that replaces the per-task MT19937 stream. */
typedef struct pv_gather_params {
    float    stepsize;
    uint32_t nused;
    float    maxdist;
    uint64_t seed;
    uint64_t ray_index_base;   /* global index of rays[0]; keeps RNG streams
                                  identical however the image is sharded        */
    uint32_t flags;            /* PV_GATHER_* */
} pv_gather_params;
#define PV_GATHER_NO_DIRECT   1u   /* skip the single-scattering term           */
#define PV_GATHER_NO_INDIRECT 2u   /* skip LPhoton                               */
/* Scheduling of the lookups (default: chosen from the search radius and the number of rays).
 *   cell-batched   every march step of the slice is sorted by photon-grid cell and one warp serves 32 neighbouring steps from
 *                  one staged block of cells; the default in the fixed-radius regime (maxdist <= cell size);
 *   step-parallel  one warp per march STEP followed by a recurrence pass (k-nearest regime, small batches of secondary rays);
 *   ray-parallel   one warp per RAY, lookups fused with the recurrence (k-nearest regime, frames of camera rays).
 * The two warp forms are bit-identical to each other; the cell-batched form sums a step's photons in photon order and agrees
 * with them to rounding (1e-6 relative).  Each form's own result does not depend on how the rays are sharded.               */
#define PV_GATHER_RAY_PARALLEL  4u
#define PV_GATHER_STEP_PARALLEL 8u
#define PV_GATHER_CELL_BATCHED  16u

/* PhotonShooter params (core/photonshooter.cpp:529-548), volume branch. */
typedef struct pv_shoot_params {
    float    stepsize;             /* shooter march step (SurfaceIntegrator "stepsize") */
    float    integrator_stepsize;  /* PhotonVolumeIntegrator stepSize (x4 in Transmittance) */
    int32_t  max_photon_depth;
    uint64_t seed;
    uint32_t rank, world;          /* emission sharding: this rank takes 4096-path
                                      blocks b with (b-1) % world == rank          */
    uint64_t max_paths;            /* safety cap on global light paths (0 = 2^40) */
    float    time;                 /* camera shutterOpen                           */
} pv_shoot_params;

typedef struct pv_shoot_stats {
    uint64_t paths;          /* global nshot when the target was met             */
    uint64_t paths_local;    /* light paths this rank traced                     */
    uint64_t photons_local;  /* volume photons this rank kept                    */
    uint64_t blocks;         /* global number of 4096-path blocks used           */
    uint64_t nodes_visited, tri_tests, density_samples, segments;
    uint64_t stack_overflows;
    double   seconds;        /* device time of the trace kernels                 */
} pv_shoot_stats;

typedef struct pv_gather_stats {
    uint64_t rays, lookups, photons_found, candidates_tested, heap_lookups;
    uint64_t shadow_rays, density_samples;
} pv_gather_stats;

/* ---- lifecycle --------------------------------------------------------- */
int         pv_create(pv_ctx **out, int device);
void        pv_destroy(pv_ctx *ctx);
const char *pv_last_error(pv_ctx *ctx);          /* ctx may be NULL             */
int         pv_version(void);

/* ---- scene (replaces the const Scene* the reference passes around) ------ */
int pv_set_scene(pv_ctx *ctx, const pv_scene_desc *scene);

/* ---- photon map: KdTree<Photon> ctor (core/kdtree.h:99-147) ------------- */
/* Inject a photon set (golden sets, checkpoints, the allgathered map).
 * SoA planes: pos[3n], wi[3n], alpha[30n].  Photon i keeps index i.          */
int pv_set_photons(pv_ctx *ctx, const float *pos, const float *wi,
                   const float *alpha, uint64_t n);
int pv_set_photons_dev(pv_ctx *ctx, const float *pos, const float *wi,
                       const float *alpha, uint64_t n);
int pv_get_photons(pv_ctx *ctx, float *pos, float *wi, float *alpha,
                   uint64_t *ids, uint64_t capacity, uint64_t *n);
int pv_get_photons_dev(pv_ctx *ctx, float *pos, float *wi, float *alpha,
                       uint64_t *ids, uint64_t capacity, uint64_t *n);
int pv_photon_count(pv_ctx *ctx, uint64_t *n);
/* Sort the photons into the uniform grid the gather walks.  maxdist / nused
 * are the integrator's lookup parameters (integrators/photonvolume.h:17-20);
 * they only size the cells: h = min(maxdist, radius expected to hold nused
 * photons).  Lookups with other parameters stay exact, just slower.         */
int pv_build(pv_ctx *ctx, float maxdist, uint32_t nused);

/* ---- KdTree::Lookup + PhotonProcess (core/kdtree.h:150-183,
 *      core/photonshooter.h:186-203) --------------------------------------- */
/* For each query point: the photons with dist^2 < r2, or if more than k of
 * them the k smallest by (dist^2, photon index).  idx/d2 are n*k, ascending
 * by (d2, idx), padded with 0xFFFFFFFF / +inf.                               */
int pv_knn(pv_ctx *ctx, const float *pts, uint64_t n, uint32_t k, float r2,
           uint32_t *idx, float *d2, uint32_t *nfound);

/* ---- Scene::Intersect / IntersectP (accelerators/bvh.cpp:585-685,
 *      shapes/trianglemesh.cpp:127-281) ------------------------------------ */
/* prim = index into the BVH's reordered primitive array or 0xFFFFFFFF.      */
int pv_intersect(pv_ctx *ctx, const pv_ray *rays, uint64_t n,
                 uint32_t *prim, float *t);
int pv_occluded(pv_ctx *ctx, const pv_ray *rays, uint64_t n, uint8_t *hit);

/* ---- BVHAccel::BVHAccel (accelerators/bvh.cpp:196-300: bounds ->
 *      recursiveBuild :301-556 -> flattenBVHTree :559-577), built on the
 *      device as a Morton-code LBVH -------------------------- */
/* prim_bounds: 6 floats per primitive (WorldBound(): pMin.xyz, pMax.xyz).
 * nodes: the reference's depth-first LinearBVHNode array, nodes_cap >=
 * 2*n_prims - 1 always suffices; *n_nodes = nodes written.  prim_order[i] =
 * index of the caller's primitive that sits at position i of the reordered
 * primitive array the leaves point into (the caller permutes tri_verts /
 * prim_material / prim_shape by it before pv_set_scene, as BVHAccel swaps
 * its `primitives` with orderedPrims, bvh.cpp:287).  A subtree of at most
 * max_prims_in_node (1..255) primitives becomes one leaf.  device_ms
 * (optional): time of the build's kernels.                                   */
int pv_build_bvh(pv_ctx *ctx, const float *prim_bounds, uint32_t n_prims,
                 uint32_t max_prims_in_node, pv_bvh_node *nodes,
                 uint32_t nodes_cap, uint32_t *n_nodes, uint32_t *prim_order,
                 float *device_ms);

/* ---- PhotonVolumeIntegrator::Transmittance (photonvolume.cpp:15-30) ----- */
/* step is the tau() step (stepSize or 4*stepSize), offset_u[n] the jitter.  */
int pv_transmittance(pv_ctx *ctx, const pv_ray *rays, uint64_t n, float step,
                     const float *offset_u, float *T);

/* ---- PhotonVolumeIntegrator::Li (photonvolume.cpp:112-222) -------------- */
/* L[30n], T[30n].                                                           */
int pv_gather(pv_ctx *ctx, const pv_ray *rays, uint64_t n,
              const pv_gather_params *params, float *L, float *T);
int pv_gather_dev(pv_ctx *ctx, const pv_ray *rays, uint64_t n,
                  const pv_gather_params *params, float *L, float *T);
/* LPhoton only (photonvolume.cpp:65-108): pts[3n], w[3n] -> L[30n].         */
int pv_lphoton(pv_ctx *ctx, const float *pts, const float *w, uint64_t n,
               uint32_t nused, float maxdist, float *L);
int pv_gather_stats_get(pv_ctx *ctx, pv_gather_stats *out, int reset);
/* Device times (ms) of the last pv_gather / pv_gather_dev call, CUDA events on its
 * stream: the gather kernel (lookups + estimate + recurrence) and the march
 * kernels that precede it (per-step medium / shadow-ray work).                */
int pv_last_kernel_ms(pv_ctx *ctx, float *ms);
int pv_last_march_ms(pv_ctx *ctx, float *ms);
/* Cell-batched schedule only, device times (ms) of the last call's parts: ms[0] sorting the march steps by cell, ms[1]
 * cellgather_kernel (lookups + flux sums: the dominant kernel), ms[2] the overflow pass (steps with more than nused photons
 * in range), ms[3] the recurrence pass.  All zero when another schedule ran.                                                */
int pv_last_phase_ms(pv_ctx *ctx, float ms[4]);
/* Kernels this context has launched so far on the map-build and gather paths (pv_build, pv_gather*): every launch site counts
 * itself; the difference across a timed region is what bench.py reports as gpu_launches.                                   */
int pv_launch_count(pv_ctx *ctx, uint64_t *n);

/* ---- SingleScatteringIntegrator::Li (integrators/single.cpp:66-138) and
 *      EmissionIntegrator::Li (integrators/emission.cpp:63-106) -------------
 * The reference's other two VolumeIntegrator plugins: the
 * same ray march with a cumulative transmittance, emission, and -- "single"
 * -- one light's single-scattered radiance per step.  No photon map is read.
 * Of params only stepsize, seed and ray_index_base are used.  L[30n], T[30n],
 * host pointers (pv_volume_li) or device pointers (pv_volume_li_dev).         */
#define PV_VOLINT_SINGLE   0
#define PV_VOLINT_EMISSION 1
/* Scheduling (params->flags; results are bit-identical either way; default: chosen from the number of rays): one warp per
 * ray with lane == spectral bin (small batches) or one thread per ray with the spectrum in registers (frames).            */
#define PV_VOLINT_WARP_PER_RAY   4u
#define PV_VOLINT_THREAD_PER_RAY 8u
int pv_volume_li(pv_ctx *ctx, int integrator, const pv_ray *rays, uint64_t n,
                 const pv_gather_params *params, float *L, float *T);
int pv_volume_li_dev(pv_ctx *ctx, int integrator, const pv_ray *rays, uint64_t n,
                     const pv_gather_params *params, float *L, float *T);

/* ---- PhotonShooter::Preprocess, volume branch (photonshooter.cpp:457-526) */
int pv_shoot(pv_ctx *ctx, uint64_t n_volume_wanted,
             const pv_shoot_params *params, pv_shoot_stats *stats);

/* Multi-rank form of the same pass.  Light paths are grouped in the reference's
 * blocks of 4096 (photonshooter.cpp:247); block b (1-based, global) holds Halton
 * indices (b-1)*4096+1 .. b*4096 and its deposits are divided by nshot = 4096*b
 * (:301,333), so the photon set does not depend on how blocks are dealt to ranks.
 * pv_shoot_blocks traces the blocks of [first_block, first_block+n_blocks) that
 * belong to params->rank and returns the number of photons each of them deposited
 * (0 for blocks of other ranks) in counts[n_blocks]; the caller sums counts over
 * ranks, finds the first block at which the running total reaches the target and
 * calls pv_shoot_finish(last_block): photons of later blocks are dropped and the
 * rest is ordered by (path index, deposit ordinal).  first_block == 1 starts a
 * new pass (clears the photon set).                                             */
int pv_shoot_blocks(pv_ctx *ctx, uint64_t first_block, uint32_t n_blocks,
                    const pv_shoot_params *params, uint32_t *counts,
                    pv_shoot_stats *stats);
int pv_shoot_finish(pv_ctx *ctx, uint64_t last_block);

/* ---- PhotonShooter::Preprocess with the SURFACE photon maps on ----------
 * followPhoton's surface branch (core/photonshooter.cpp:147-189): caustic / indirect / direct
 * deposits and radiance-photon sites, the per-task done flags that change path behaviour at
 * block ends (:239-241, :303-341), the give-up rule over all wanted counts (:285-299), then
 * ComputeRadianceTask + EPhoton (:17-35, :359-395).  The reference's PhotonIntegrator keeps
 * reading these maps as KdTree<Photon> / KdTree<RadiancePhoton> (integrators/photonmap.cpp:
 * 159-164); the adapter fills them from pv_get_map_photons / pv_radiance_photons.            */
enum { PV_MAP_VOLUME = 0, PV_MAP_CAUSTIC = 1, PV_MAP_INDIRECT = 2, PV_MAP_DIRECT = 3,
       PV_MAP_RADIANCE = 4 };
typedef struct pv_maps_params {
    uint64_t n_volume_wanted, n_caustic_wanted, n_indirect_wanted;   /* photonshooter.cpp:529-533 */
    int32_t  final_gather;       /* direct photons and radiance photons only with final gathering */
} pv_maps_params;
typedef struct pv_maps_stats {
    uint64_t nshot, blocks;      /* light paths / 4096-path blocks of the whole pass               */
    uint64_t n_caustic_paths, n_indirect_paths, n_direct_paths, n_volume_paths;  /* the divisors the
                                    estimates use (:305,:313,:321,:331; n_volume_paths includes :104) */
    uint64_t n[5];               /* photons kept per PV_MAP_* class                                */
    uint64_t replayed_blocks;    /* blocks traced twice because a done flag flipped inside a wave   */
    pv_shoot_stats shoot;
} pv_maps_stats;
/* One pass, single rank.  Afterwards the context's photon set (pv_get_photons, pv_build) is the
 * volume map; the other classes are read with pv_get_map_photons.  Photons of every class are
 * ordered by (light path, deposit ordinal): the order one reference task stores them in.
 * ids = class << 60 | path index << 16 | deposit ordinal along the path.                         */
int pv_shoot_maps(pv_ctx *ctx, const pv_maps_params *maps, const pv_shoot_params *params,
                  pv_maps_stats *stats);
/* The same pass sharded over params->world ranks (one process per GPU): rank r traces the 4096-path
 * blocks b with (b-1) % world == r.  `allreduce` must sum data[0..n) in place over all ranks
 * (ncclAllReduce / torch.distributed.all_reduce on a staging tensor) and return 0; it is called once
 * per wave with the per-class, per-block deposit counts, after which every rank replays the same
 * bookkeeping, so flags, roll-backs and the last block agree everywhere.  Afterwards each rank
 * holds ITS photons of every class (ordered by id); the union over ranks, ordered by id, is exactly
 * the single-rank result.  stats->n[] are local counts, the path counts and nshot are global.      */
typedef int (*pv_allreduce_u32_fn)(uint32_t *data, uint64_t n, void *user);
int pv_shoot_maps_ranks(pv_ctx *ctx, const pv_maps_params *maps, const pv_shoot_params *params,
                        pv_allreduce_u32_fn allreduce, void *user, pv_maps_stats *stats);
/* SoA planes like pv_get_photons.  PV_MAP_RADIANCE: wi = faceforwarded normal, alpha = rho_r
 * (rho_t == 0: only matte surfaces hold radiance photons on this path).                          */
int pv_get_map_photons(pv_ctx *ctx, int map, float *pos, float *wi, float *alpha,
                       uint64_t *ids, uint64_t capacity, uint64_t *n);
/* Inject a class (golden sets): map in PV_MAP_CAUSTIC..PV_MAP_RADIANCE.                          */
int pv_set_map_photons(pv_ctx *ctx, int map, const float *pos, const float *wi,
                       const float *alpha, uint64_t n);
/* ComputeRadianceTask::Run for every radiance photon: Lo = INV_PI * rho_r * (E_direct +
 * E_indirect + E_caustic), EPhoton = n_lookup nearest photons within max_dist2 facing the
 * normal, over path count * final search radius^2 * pi.  path_counts = nDirectPaths,
 * nIndirectPaths, nCausticPaths (NULL: the counts of the last pv_shoot_maps).  Builds a grid per
 * class, so the volume map must be (re)built with pv_build afterwards.  Lo[30 * capacity].       */
int pv_radiance_photons(pv_ctx *ctx, uint32_t n_lookup, float max_dist2,
                        const uint64_t *path_counts, float *Lo, uint64_t capacity, uint64_t *n);

/* Inject the radiance of the PV_MAP_RADIANCE photons (Lo[30n], n = their count) instead of computing it with
 * pv_radiance_photons: a second context that serves final-gather rays while the first one keeps its grid on the volume map. */
int pv_set_radiance_lo(pv_ctx *ctx, const float *Lo, uint64_t n);

/* ---- building blocks of the SURFACE integrator's lookups (PhotonIntegrator, integrators/photonmap.cpp) -------
 * There is one lookup grid per context.  pv_build builds it over the volume photons; pv_select_map builds it over any
 * photon class (PV_MAP_*), after which pv_knn and the two calls below query that class.  pv_gather / pv_lphoton refuse to
 * run until pv_build has put the volume map back.                                                                  */
int pv_select_map(pv_ctx *ctx, int map, float maxdist, uint32_t nused);
/* LPhoton, diffuse branch (integrators/photonmap.cpp:62-108 with kernel() :57-60) on the selected caustic / indirect /
 * direct map: pts[3n], nf[3n] = Faceforward(shading normal, wo).  Lr[30n] / Lt[30n] = kernel-weighted flux of the
 * n_lookup nearest photons within max_dist2 arriving on the side of nf / on the other side, over n_paths * md2; the
 * caller finishes with L = Lr * rho_r / pi + Lt * rho_t / pi.                                                      */
int pv_surface_lphoton(pv_ctx *ctx, const float *pts, const float *nf, uint64_t n, uint32_t n_lookup,
                       float max_dist2, uint64_t n_paths, float *Lr, float *Lt);
/* RadiancePhotonProcess + KdTree::Lookup with an unbounded radius (core/photonshooter.h:54-70, final gathering
 * integrators/photonmap.cpp:238-243) on the selected radiance-photon map: for each (point, normal) the nearest
 * radiance photon whose normal has a positive dot product with it.  idx[n] = its index in the PV_MAP_RADIANCE list
 * or 0xFFFFFFFF; Lo[30n] (may be NULL) = its radiance from the last pv_radiance_photons.                           */
int pv_radiance_nearest(pv_ctx *ctx, const float *pts, const float *normals, uint64_t n, uint32_t *idx, float *Lo);
/* One batch of final-gather rays (integrators/photonmap.cpp:231-243, :278-289), the radiance-photon map selected: each ray is
 * traced (Scene::Intersect), the nearest radiance photon facing Faceforward(hit normal, -d) is looked up at the hit, and its
 * Lo is attenuated by renderer->Transmittance(ray, sample == NULL) = exp(-tau(ray up to the hit, step, u)) with
 * step = 4 * the volume integrator's stepsize and u a Philox draw keyed by (seed, index_base + i).  Lindir[30n]; 0 for rays
 * that hit nothing or find no facing photon.  idx[n] (may be NULL) = the radiance photon used or 0xFFFFFFFF.                */
int pv_final_gather(pv_ctx *ctx, const pv_ray *rays, uint64_t n, float step, uint64_t seed, uint64_t index_base,
                    float *Lindir, uint32_t *idx);

/* ---- multi-GPU: replication of the photon map over NVLink ------------------------------------------
 * The reference's photon map is one vector shared by its pthread tasks (core/photonshooter.cpp:333-337); with emission
 * sharded over GPUs (pv_shoot_blocks) the map is replicated once per frame instead.  NCCL is opened at run time
 * (libnccl.so.2); a single-GPU host never needs it.
 *   one process per GPU: rank 0 calls pv_comm_unique_id and hands the 128 bytes to the others (MPI, a file, a TCP store),
 *     every rank calls pv_comm_init, then pv_allgather_photons after shooting (or after pv_set_photons of its slice):
 *     every rank ends with the union of all ranks' photons ordered by id -- for shot photons (renumber = 0) exactly the
 *     single-rank set; renumber = 1 first gives rank r's photon i the id (photons of ranks < r) + i, i.e. the index it has
 *     in the concatenated set (injected slices).  collective_ms (may be NULL): device time of the grouped all-gathers.
 *   one process, several GPUs (the drop-in's PV_DEVICES): pv_comm_init_all over one context per device, then
 *     pv_broadcast_photons sends ctxs[src]'s photon set to the others.                                               */
#define PV_COMM_ID_BYTES 128
int pv_comm_unique_id(uint8_t *id);
int pv_comm_init(pv_ctx *ctx, const uint8_t *id, int rank, int world);
int pv_comm_init_all(pv_ctx **ctxs, int n);
int pv_comm_destroy(pv_ctx *ctx);
int pv_allgather_photons(pv_ctx *ctx, int renumber, float *collective_ms);
int pv_broadcast_photons(pv_ctx **ctxs, int n, int src, float *collective_ms);

/* Li with a stream index of its own per ray: ray i draws from the stream ray_index[i] (params->ray_index_base is not read), so the
 * result of a ray does not depend on which other rays share the call.  For callers that collect rays from many threads -- the rays
 * specular bounces spawn deep inside the reference's recursive surface shading (SpecularReflect -> Renderer::Li ->
 * VolumeIntegrator::Li, core/integrator.cpp:280-340) reach the volume integrator one at a time on every render thread.  Host
 * pointers.  pv_volume_li_indexed: integrator = PV_VOLINT_*.                                                                */
int pv_gather_indexed(pv_ctx *ctx, const pv_ray *rays, const uint64_t *ray_index, uint64_t n, const pv_gather_params *params,
                      float *L, float *T);
int pv_volume_li_indexed(pv_ctx *ctx, int integrator, const pv_ray *rays, const uint64_t *ray_index, uint64_t n,
                         const pv_gather_params *params, float *L, float *T);

/* raw CUDA stream (cudaStream_t) the context launches on, for event timing */
void *pv_stream(pv_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PV_H */
