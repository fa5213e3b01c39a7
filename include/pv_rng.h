/* pv_rng.h -- counter-based random streams of the B200 photon-volume path.
 *
 * The reference draws everything that is not a Halton dimension from one
 * sequential MT19937 stream per pthread task (core/rng.cpp:43-107;
 * RNG(31*taskNum) core/photonshooter.cpp:235, RNG(taskNum)
 * renderers/samplerrenderer.cpp:73).  A sequential stream cannot be replayed
 * by 10^5 GPU threads, so this path keys Philox4x32-10 by what the draw is
 * FOR; the mapping below is the whole definition and is shared by the CUDA
 * kernels and by the CPU oracle so both consume identical numbers.
 *
 *  gather, per ray   : ctr = (ray_lo, ray_hi, 0,    PV_RNG_RAY ) -> w0 = VdC scramble for
 *                      the light-choice sequence, w1 = its permutation key
 *  gather, per step  : ctr = (ray_lo, ray_hi, step, PV_RNG_STEP) -> w0 = tau offset
 *                      (photonvolume.cpp:154), w1 = roulette (:160), w2 = shadow-ray
 *                      tau offset (:195 -> :26)
 *  shooter, per path : ctr = (path_lo, path_hi, j,  PV_RNG_PATH) -> words 4j..4j+3 of the
 *                      path's stream, consumed in the reference's draw order
 *  key               = (seed_lo, seed_hi)
 *
 * RandomFloat keeps the reference's mapping (core/rng.cpp:59-65):
 * (u32 & 0xffffff) / 2^24.
 */
#ifndef PV_RNG_H
#define PV_RNG_H
#include <stdint.h>

#if defined(__CUDACC__)
#define PV_HD __host__ __device__ __forceinline__
#else
#define PV_HD static inline
#endif

#define PV_RNG_RAY  0x7261u
#define PV_RNG_STEP 0x6761u
#define PV_RNG_PATH 0x7368u
#define PV_RNG_AREA 0x6172u           /* area-light sample of a march step: ctr = (ray_lo, ray_hi, step, tag) -> w0, w1 = position on the light, w2 = component */
#define PV_RNG_FINAL_GATHER 0x6667u   /* final-gather ray: ctr = (ray_lo, ray_hi, 0, tag) -> w0 = transmittance offset */

PV_HD void pv_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                            uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

PV_HD float pv_u32_to_float(uint32_t u) {
    return (float)(u & 0xffffffu) / 16777216.0f;
}

/* Stateless random permutation of [0,l) (Kensler 2013); stands in for the
 * reference's Shuffle() of the per-step light samples (core/montecarlo.h:304-312). */
PV_HD uint32_t pv_permute(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p;             i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;  i *= 1u | p >> 27;
                            i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2;  i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;  i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}

/* core/montecarlo.h:277-286 */
PV_HD float pv_van_der_corput(uint32_t n, uint32_t scramble) {
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
    n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
    n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
    n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
    n ^= scramble;
    float v = (float)((n >> 8) & 0xffffffu) / 16777216.0f;
    const float one_minus_eps = 0.99999994f; /* 0x1.fffffep-1 */
    return v < one_minus_eps ? v : one_minus_eps;
}

#endif
