/* pv_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's volumetric photon-mapping path
 * (piwell/CS348B-pbrt).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the reported CPU baseline -- never as the product path.
 *
 * Parity status: PINNED.  tests/golden/ holds vectors produced by the real
 * reference (oracle/_ref/ref_harness, built from /root/reference by
 * oracle/Makefile); tests/test_oracle_golden.py checks every function below
 * against them.
 */
#ifndef PV_ORACLE_H
#define PV_ORACLE_H
#include <stdint.h>
#include "../include/pv.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvo_kdtree pvo_kdtree;

/* RNG streams.  PVO_RNG_PHILOX is the counter-based stream the CUDA path uses
 * (DESIGN.md "Random streams"); PVO_RNG_MT replays the reference's MT19937
 * order so results can be compared with the reference binary itself. */
#define PVO_RNG_PHILOX 0
#define PVO_RNG_MT     1

/* core/kdtree.h:99-147 */
pvo_kdtree *pvo_kdtree_build(const float *pos, uint64_t n);
void        pvo_kdtree_free(pvo_kdtree *t);
/* core/kdtree.h:150-183 + core/photonshooter.h:186-203; output ascending by
 * (d2, original photon index), padded with 0xFFFFFFFF / +inf.
 * boundary_ties (may be NULL) counts queries whose k-th distance is shared
 * with a rejected candidate, i.e. where the reference's result is
 * traversal-order dependent. */
int pvo_knn(const pvo_kdtree *t, const float *pts, uint64_t n, uint32_t k,
            float r2, uint32_t *idx, float *d2, uint32_t *nfound,
            uint64_t *boundary_ties);
/* brute force: k smallest by (d2, index) among d2 < r2 -- the rule the CUDA
 * path implements. */
int pvo_knn_brute(const float *pos, uint64_t nph, const float *pts, uint64_t n,
                  uint32_t k, float r2, uint32_t *idx, float *d2, uint32_t *nfound);

/* integrators/photonvolume.cpp:65-108 */
int pvo_lphoton(const pv_scene_desc *sc, const pvo_kdtree *t, const float *wi,
                const float *alpha, const float *pts, const float *w, uint64_t n,
                uint32_t nused, float maxdist, float *L);

/* accelerators/bvh.cpp:585-685, shapes/trianglemesh.cpp:127-281 */
int pvo_intersect(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n,
                  uint32_t *prim, float *t);
int pvo_occluded(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n, uint8_t *hit);

/* integrators/photonvolume.cpp:15-30 with explicit step and offset */
int pvo_transmittance(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n,
                      float step, const float *offset_u, float *T);

/* integrators/photonvolume.cpp:112-222.  rng_mode PVO_RNG_MT seeds
 * RNG(mt_seed + i) for ray i exactly like oracle/ref_harness.cpp --li. */
int pvo_gather(const pv_scene_desc *sc, const pvo_kdtree *t, const float *wi,
               const float *alpha, const pv_ray *rays, uint64_t n,
               const pv_gather_params *prm, int rng_mode, uint32_t mt_seed,
               int nthreads, float *L, float *T, pv_gather_stats *stats);

/* AggregateVolume (core/volume.cpp:178-261; what the reference builds for a scene with several Volume statements).
 * pv_scene_desc carries one medium -- the device path has no aggregate yet -- so the regions after the
 * first are handed to the oracle on the side; every function below then sees the aggregate.  (NULL, 0) switches it off.
 * Not thread-safe with respect to the other calls: set it, call, reset it. */
void pvo_set_more_media(const pv_medium *more, uint32_t n);

/* DiffuseAreaLight over a triangle mesh (lights/diffuse.cpp, ShapeSet core/light.cpp:114-172), groundwork for DESIGN.md 11.2.
 * pv_light has no area-light fields yet: the light keeps its slot in sc->lights[] as a placeholder of type PVO_LIGHT_SLOT
 * (what oracle/ref_harness --export-area-lights writes) and its data is handed over here.  Covered: the point-query
 * Sample_L of the volume integrators' direct term, with the reference's MT stream. */
#define PVO_LIGHT_SLOT 100
typedef struct pvo_area_light {
    uint32_t slot;            /* index in the scene's light list */
    uint32_t n_tris;
    uint32_t flags;           /* 1 = ReverseOrientation, 2 = TransformSwapsHandedness */
    uint32_t pad;
    const float *tri;         /* 9 floats per triangle, world space, in the ShapeSet's (refine) order */
    float Lemit[30];
} pvo_area_light;
void pvo_set_area_lights(const pvo_area_light *lights, uint32_t n);

/* The reference's other two volume integrators: SingleScatteringIntegrator::Li
 * (integrators/single.cpp:66-138) and EmissionIntegrator::Li (integrators/emission.cpp:63-106).  Only
 * prm->stepsize / seed / ray_index_base are read.  PVO_RNG_MT seeds RNG(mt_seed + i) for ray i like
 * oracle/ref_harness.cpp --vli. */
#define PVO_VLI_SINGLE   0
#define PVO_VLI_EMISSION 1
int pvo_volume_li(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n, const pv_gather_params *prm, int kind,
                  int rng_mode, uint32_t mt_seed, float *L, float *T, pv_gather_stats *stats);

/* core/photonshooter.cpp:47-357, volume branch (caustic/indirect maps off).
 * PVO_RNG_MT: one task (taskNum 0), the reference's sequential stream.
 * PVO_RNG_PHILOX: per-path streams; blocks may be spread over nthreads. */
typedef struct pvo_photons {
    uint64_t  n;
    float    *pos, *wi, *alpha;   /* malloc'd: 3n, 3n, 30n */
    uint64_t *ids;                /* (path index << 16) | deposit ordinal */
    uint64_t  nshot, blocks;
    uint64_t  nodes_visited, tri_tests, density_samples, segments;
} pvo_photons;
int  pvo_shoot(const pv_scene_desc *sc, uint64_t n_wanted, const pv_shoot_params *prm,
               int rng_mode, int nthreads, pvo_photons *out);
void pvo_photons_free(pvo_photons *p);

/* The same pass with the surface maps on (photonshooter.cpp:147-189,303-341), ONE task.
 * cls[0] volume (alpha / nshot), cls[1] caustic, cls[2] indirect, cls[3] direct, cls[4] radiance-photon sites
 * (wi plane = faceforwarded normal, alpha plane = rho_r; rho_t == 0 for the matte surfaces that can hold one).
 * ids = class << 60 | path index << 16 | deposit ordinal along the path. */
typedef struct pvo_maps {
    pvo_photons cls[5];
    uint64_t nshot, blocks;
    uint64_t n_caustic_paths, n_indirect_paths, n_direct_paths, n_volume_paths;
} pvo_maps;
int  pvo_shoot_maps(const pv_scene_desc *sc, uint64_t n_volume, uint64_t n_caustic, uint64_t n_indirect, int final_gather,
                    const pv_shoot_params *prm, int rng_mode, pvo_maps *out);
void pvo_maps_free(pvo_maps *m);
/* ComputeRadianceTask::Run + EPhoton (photonshooter.cpp:17-35,359-395), rho_t == 0: maps in the order direct, indirect,
 * caustic (NULL = absent), counts = nDirectPaths, nIndirectPaths, nCausticPaths. Lo[30n]. */
int  pvo_radiance(const pvo_kdtree *maps[3], const float *wis[3], const float *alphas[3], const uint64_t counts[3],
                  const float *rp_pos, const float *rp_n, const float *rho_r, uint64_t n, uint32_t nLookup, float maxDist2, float *Lo);

/* the surface integrator's two photon lookups (integrators/photonmap.cpp:62-108 diffuse branch; :238-243) */
int  pvo_surface_lphoton(const pvo_kdtree *t, const float *wi, const float *alpha, const float *pts, const float *nf, uint64_t n,
                         uint32_t nLookup, float maxDist2, uint64_t nPaths, float *Lr, float *Lt);
int  pvo_radiance_nearest(const float *rp_pos, const float *rp_n, uint64_t n_rp, const float *pts, const float *nrm, uint64_t n,
                          uint32_t *idx, float *d2out);

int  pvo_final_gather(const pv_scene_desc *sc, const float *rp_pos, const float *rp_n, const float *rp_Lo, uint64_t n_rp, const pv_ray *rays,
                      uint64_t n, float step, uint64_t seed, uint64_t index_base, float *Lindir, uint32_t *idx);

/* small known-answer helpers exposed for unit tests */
uint32_t pvo_mt_first(uint32_t seed, uint32_t *out, uint32_t n);
void     pvo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void     pvo_halton6(uint32_t mt_seed, uint32_t n, float out[6]);
float    pvo_spectrum_y(const pv_scene_desc *sc, const float *c);
uint32_t pvo_permute(uint32_t i, uint32_t l, uint32_t p);
float    pvo_van_der_corput(uint32_t n, uint32_t scramble);

#ifdef __cplusplus
}
#endif
#endif
