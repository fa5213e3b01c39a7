// pv_shoot.cuh -- what the photon shooter's kernels share (pv_shoot.cu: the persistent-thread state-machine kernel;
// pv_wavefront.cu: the wavefront kernels): continuation frames, launch arguments, the per-path Philox stream, the
// warp-aggregated deposit and the reference's sampling routines.
#pragma once
#include <cooperative_groups.h>
#include "pv_ctx.h"
namespace cg = cooperative_groups;

#define SH_THREADS 128
#define SH_MAXDEPTH 24
#define SH_BLOCK 4096

enum { ST_NEWPATH = 0, ST_TRACE = 1, ST_SURFACE = 2, ST_DONE = 3 };

struct Frame {
    float o[3], d[3], mint, maxt;
    float ip[3], inn[3], idpdu[3], ieps;      // hit: dg.p, dg.nn, dg.dpdu, rayEpsilon
    int prim, nI, spec, loop_i;                // spec: bit 0 = specularPath; bit 1 = the fork's Spectrum::lambda >= 0, i.e. alpha had
                                               // exactly one positive bin WHEN IT WAS LAST RE-MADE (emission, surface bounce:
                                               // extractLambda in SampledSpectrum's converting constructor, core/spectrum.h:266-279,
                                               // :339-343); in-place updates in between (scatter, transmittance) carry it along
    float alpha[PV_NSPEC];
};

struct ShootArgs {
    const DevScene *sc;
    uint64_t b_start;               // first global block (1-based) of this rank in the wave
    uint32_t n_local_blocks, world;
    uint64_t first_block;           // wave origin, for block_counts indexing
    float stepsize, istep4;
    int max_depth;
    uint32_t k0, k1;
    uint32_t perm[41];              // PermutedHalton tables of task 0 (RNG(31*0)), montecarlo.cpp:380-397
    float *pos, *wi, *alpha32; uint64_t *ids;
    unsigned long long *n_out; uint64_t cap;
    uint32_t *block_counts;               // [class][block of the wave]
    uint32_t wave_blocks;                 // row length of block_counts
    uint32_t flags;                       // SF_*: the reference's per-task done flags, constant over a wave
    unsigned long long *work, *stats;     // stats: nodes, tris, density samples, segments, overflows, paths, first-hit scatters
};
// !causticDone, !indirectDone, volumeDone, finalGather (photonshooter.cpp:239-241, :411)
enum { SF_WANT_CAUSTIC = 1, SF_WANT_INDIRECT = 2, SF_VOLUME_DONE = 4, SF_FINAL_GATHER = 8 };
// photon classes: the top 4 bits of a photon id
enum { PC_VOLUME = 0, PC_CAUSTIC = 1, PC_INDIRECT = 2, PC_DIRECT = 3, PC_RADIANCE = 4, PC_COUNT = 5 };

// pv_wavefront.cu
int pvi_wavefront_run(pv_ctx *ctx, const ShootArgs &a, bool surf, int kind, bool *replay);

static __device__ __noinline__ uint4 path_philox_block(uint32_t c0, uint32_t c1, uint32_t j, uint32_t k0, uint32_t k1) {
    uint32_t out[4];
    pv_philox4x32_10(c0, c1, j, PV_RNG_PATH, k0, k1, out);
    return make_uint4(out[0], out[1], out[2], out[3]);
}
struct PathRng {
    uint32_t c0, c1, j, pos, k0, k1, buf[4];
    __device__ __forceinline__ void reset(uint64_t path, uint32_t key0, uint32_t key1) {
        c0 = (uint32_t)path; c1 = (uint32_t)(path >> 32); j = 0; pos = 4; k0 = key0; k1 = key1;
    }
    __device__ __forceinline__ float next() {
        if (pos == 4) {                  // one out-of-line Philox block per four draws: ~25 call sites would otherwise inline 10 rounds each
            const uint4 r = path_philox_block(c0, c1, j++, k0, k1);
            buf[0] = r.x; buf[1] = r.y; buf[2] = r.z; buf[3] = r.w; pos = 0;
        }
        uint32_t v = pos == 0 ? buf[0] : (pos == 1 ? buf[1] : (pos == 2 ? buf[2] : buf[3]));
        pos++;
        return pv_u32_to_float(v);
    }
    // discard n draws (BSDF::rho's stratified samples, which a Lambertian BRDF never reads)
    __device__ __forceinline__ void skip(uint32_t n) {
        const uint32_t consumed = j * 4 - (4 - pos) + n;
        const uint32_t q = consumed >> 2, r = consumed & 3;
        if (r == 0) { j = q; pos = 4; }
        else { const uint4 b = path_philox_block(c0, c1, q, k0, k1); buf[0] = b.x; buf[1] = b.y; buf[2] = b.z; buf[3] = b.w; j = q + 1; pos = r; }
    }
};

// Append one photon of class `cls`: one atomicAdd per coalesced group (warp-aggregated), alpha as one 128-byte line.
// Volume photons are divided by nshot of their block at deposit time (photonshooter.cpp:333); surface photons are not.
__device__ __forceinline__ void deposit_photon(const ShootArgs &a, uint32_t cls, uint64_t gblock, uint64_t path, uint32_t dep_seq, v3 p, v3 w,
                                               const float *alpha, float fn) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long slot = 0;
    if (g.thread_rank() == 0) slot = atomicAdd(a.n_out, (unsigned long long)g.size());
    slot = g.shfl(slot, 0) + g.thread_rank();
    atomicAdd(&a.block_counts[cls * a.wave_blocks + (uint32_t)(gblock - a.first_block)], 1u);
    if (slot < a.cap) {
        a.pos[3 * slot] = p.x; a.pos[3 * slot + 1] = p.y; a.pos[3 * slot + 2] = p.z;
        a.wi[3 * slot] = w.x; a.wi[3 * slot + 1] = w.y; a.wi[3 * slot + 2] = w.z;
        float4 *dst = reinterpret_cast<float4 *>(a.alpha32 + 32 * slot);
#pragma unroll 1
        for (int q = 0; q < 7; ++q)
            dst[q] = make_float4(__fdiv_rn(alpha[4 * q], fn), __fdiv_rn(alpha[4 * q + 1], fn),
                                 __fdiv_rn(alpha[4 * q + 2], fn), __fdiv_rn(alpha[4 * q + 3], fn));
        dst[7] = make_float4(__fdiv_rn(alpha[28], fn), __fdiv_rn(alpha[29], fn), 0.f, 0.f);
        a.ids[slot] = ((uint64_t)cls << 60) | (path << 16) | (uint64_t)(dep_seq & 0xffffu);
    }
}

__device__ __forceinline__ v3 uniform_sample_sphere(float u1, float u2) {       // core/montecarlo.cpp:283-290
    float z = 1.f - 2.f * u1;
    float r = __fsqrt_rn(fmaxf(0.f, 1.f - z * z));
    float phi = 2.f * PV_PI_F * u2, sp, cp;
    sincosf(phi, &sp, &cp);                       // one argument reduction for both
    return V3(r * cp, r * sp, z);
}
__device__ __forceinline__ v3 uniform_sample_cone(float u1, float u2, float costhetamax) {   // :405-410
    float costheta = (1.f - u1) + u1 * costhetamax;
    float sintheta = __fsqrt_rn(1.f - costheta * costheta);
    float phi = u2 * 2.f * PV_PI_F, sp, cp;
    sincosf(phi, &sp, &cp);
    return V3(cp * sintheta, sp * sintheta, costheta);
}
__device__ __forceinline__ void concentric_sample_disk(float u1, float u2, float *dx, float *dy) {   // :306-348
    float r, theta;
    float sx = 2 * u1 - 1, sy = 2 * u2 - 1;
    if (sx == 0.f && sy == 0.f) { *dx = 0.f; *dy = 0.f; return; }
    if (sx >= -sy) {
        if (sx > sy) { r = sx; if (sy > 0.f) theta = __fdiv_rn(sy, r); else theta = 8.0f + __fdiv_rn(sy, r); }
        else { r = sy; theta = 2.0f - __fdiv_rn(sx, r); }
    } else {
        if (sx <= sy) { r = -sx; theta = 4.0f - __fdiv_rn(sy, r); }
        else { r = -sy; theta = 6.0f + __fdiv_rn(sx, r); }
    }
    theta *= PV_PI_F / 4.f;
    float st, ct;
    sincosf(theta, &st, &ct);
    *dx = r * ct;
    *dy = r * st;
}
__device__ __forceinline__ void coordinate_system(v3 v1, v3 *v2, v3 *v3o) {      // core/geometry.h:508-518
    if (fabsf(v1.x) > fabsf(v1.y)) {
        float invLen = __fdiv_rn(1.f, __fsqrt_rn(v1.x * v1.x + v1.z * v1.z));
        *v2 = V3(-v1.z * invLen, 0.f, v1.x * invLen);
    } else {
        float invLen = __fdiv_rn(1.f, __fsqrt_rn(v1.y * v1.y + v1.z * v1.z));
        *v2 = V3(0.f, v1.z * invLen, -v1.y * invLen);
    }
    *v3o = vcross(v1, *v2);
}
// FresnelDielectric::Evaluate + FrDiel (core/reflection.cpp:60-67,115-135), eta_i = 1, eta_t = ior
__device__ __forceinline__ float fresnel_dielectric(float cosi, float ior) {
    cosi = fminf(fmaxf(cosi, -1.f), 1.f);
    bool entering = cosi > 0.f;
    float ei = 1.f, et = ior;
    if (!entering) { float t = ei; ei = et; et = t; }
    float sint = __fdiv_rn(ei, et) * __fsqrt_rn(fmaxf(0.f, 1.f - cosi * cosi));
    if (sint >= 1.f) return 1.f;
    float cost = __fsqrt_rn(fmaxf(0.f, 1.f - sint * sint));
    float ac = fabsf(cosi);
    float Rparl = __fdiv_rn((et * ac) - (ei * cost), (et * ac) + (ei * cost));
    float Rperp = __fdiv_rn((ei * ac) - (et * cost), (ei * ac) + (et * cost));
    return __fdiv_rn(Rparl * Rparl + Rperp * Rperp, 2.f);
}
