// pv_volint.cu -- the reference's other two volume integrators on the march records:
//   SingleScatteringIntegrator::Li   integrators/single.cpp:66-138
//   EmissionIntegrator::Li           integrators/emission.cpp:63-106
//
// Both walk the camera ray exactly like PhotonVolumeIntegrator::Li (same nSamples / step / jittered start / step-segment
// optical depth / light choice / shadow ray), so the march kernels of pv_march.cu produce their per-step scalars unchanged
// ("emission" runs them with the direct term off).  What differs is the spectral recurrence, done here with one warp per
// ray and lane == spectral bin:
//   * the transmittance is CUMULATIVE: Tr *= Exp(-stepTau) (single.cpp:100, emission.cpp:89) -- the photon-volume integrator
//     of this fork restarts it every step (photonvolume.cpp:155);
//   * Russian roulette is decided on that cumulative Tr.y() (single.cpp:103-110), so the draw cannot be made by the march
//     kernel: it is re-derived here from the same keyed Philox word (PV_RNG_STEP, word 1) when the test fires;
//   * per step Lv += Tr * Lve(p) and, "single" only, Lv += Tr * ss * p(p, w, -wo) * Ld * nLights / pdf (single.cpp:113-130);
//     the result is Lv * step, *T = Tr.
// Two schedules of the same arithmetic (bit-identical results, tests/test_gpu_volint.py):
//   volint_kernel         one WARP per ray, lane == bin.  Lowest latency per ray; but the per-step bookkeeping (five shuffles,
//                         loop control) is paid by all 32 lanes, and ncu shows the kernel issue-bound (76 % issue slots busy,
//                         72 warp instructions per step, 4.5 ms for a 1080p frame of "single").  Used for small batches.
//   volint_thread_kernel  one THREAD per ray, the 30 bins of Tr and Lv in registers, medium / light spectra in shared
//                         memory.  The bookkeeping is paid once per ray-step instead of once per 32 lanes.  Used for frames.
// Streaming kernels: 32 B of records in per step, 240 B out per ray.
#include <algorithm>
#include "pv_ctx.h"
#include "pv_march.cuh"

#define VI_THREADS 128

struct VolIntArgs {
    const DevScene *sc;
    const pv_ray *rays;
    const RayHdr *hdr;
    const StepRec *steps;
    uint64_t n;
    uint32_t k0, k1;               // Philox key
    uint64_t ray_index_base;
    const uint64_t *ray_index;     // per-ray stream indices (pv_volume_li_indexed) or null
    float *L, *T;
};

// wo of Light::Sample_L(p, ...) (lights/point.cpp:50-57, spot.cpp:50-57, distant.cpp:48-55)
static __device__ __forceinline__ v3 vi_light_wo(const pv_light &l, v3 p) {
    if (l.type == PV_LIGHT_DISTANT) return V3(l.dir[0], l.dir[1], l.dir[2]);
    return vnorm(V3(l.pos[0], l.pos[1], l.pos[2]) - p);
}

template <bool SINGLE>
__global__ void __launch_bounds__(VI_THREADS) volint_kernel(VolIntArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t ri = ((uint64_t)blockIdx.x * VI_THREADS + threadIdx.x) >> 5;      // one warp per ray
    if (ri >= a.n) return;
    const DevScene &sc = *a.sc;
    const DevMedium &med = sc.med;
    const bool bin = lane < PV_NSPEC;
    const float sig_a = bin ? med.sigma_a[lane] : 0.f, sig_s = bin ? med.sigma_s[lane] : 0.f, le = bin ? med.le[lane] : 0.f;
    const float sig_t = sig_a + sig_s;
    const float cy = bin ? sc.cie_y[lane] : 0.f;
    float sig_t_max = sig_t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sig_t_max = fmaxf(sig_t_max, __shfl_xor_sync(PV_FULL, sig_t_max, o));
    const bool rainbow = med.type == PV_MEDIUM_RAINBOW;
    const int nLights = (int)sc.n_lights;

    const float4 h0 = __ldg(reinterpret_cast<const float4 *>(a.hdr + ri));
    const int nSamples = __float_as_int(h0.z);
    const float step = h0.w;
    const StepRec *recs = a.steps + (((unsigned long long)__float_as_uint(h0.y) << 32) | __float_as_uint(h0.x));
    float Tr = 1.f, Lv = 0.f;
    if (nSamples > 0) {
        const v3 ro = V3(__ldg(&a.rays[ri].o[0]), __ldg(&a.rays[ri].o[1]), __ldg(&a.rays[ri].o[2]));
        const v3 rd = V3(__ldg(&a.rays[ri].d[0]), __ldg(&a.rays[ri].d[1]), __ldg(&a.rays[ri].d[2]));
        const uint64_t gidx = a.ray_index ? a.ray_index[ri] : a.ray_index_base + ri;
        float S = 0.f;                                      // sum of the step optical-depth scalars: Tr[b] >= exp(-sig_t[b] * S)
        bool stop = false;
        for (int c0 = 0; c0 < nSamples && !stop; c0 += 32) {
            // lane == step: the march records of the next 32 steps
            float c_t = 0.f, c_tau = 0.f, c_dens = 0.f, c_sh = 0.f, c_dfac = 0.f;
            int c_ln = 0;
            if (c0 + (int)lane < nSamples) {
                const float4 *rp = reinterpret_cast<const float4 *>(recs + c0 + lane);
                const float4 ra = __ldg(rp), rb = __ldg(rp + 1);
                c_t = ra.x; c_tau = ra.y; c_dens = ra.w; c_sh = rb.x; c_dfac = rb.y; c_ln = __float_as_int(rb.z);
            }
            // lane == bin: the recurrence, one step at a time
            const int nthis = min(32, nSamples - c0);
            for (int i = 0; i < nthis; ++i) {
                const float s_tau = __shfl_sync(PV_FULL, c_tau, i);
                Tr = Tr * expf(-(sig_t * s_tau));                          // Tr *= Exp(-stepTau)
                S += s_tau;
                // Tr.y() < 1e-3 ?  y(1) ~ 1 and every bin is >= exp(-sig_t_max * S) (the roulette only ever doubles Tr), so the
                // 30-term sum is needed only once the bound itself is small
                if (sig_t_max * S > 6.0f) {
                    float yy = 0.f;                                        // core/spectrum.h:433-439: bins summed in order
                    for (int bb = 0; bb < PV_NSPEC; ++bb) yy += __shfl_sync(PV_FULL, cy, bb) * __shfl_sync(PV_FULL, Tr, bb);
                    if (__fdiv_rn(yy * 300.f, 106.856895f * (float)PV_NSPEC) < 1e-3f) {
                        uint32_t sw[4];
                        pv_philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)(c0 + i), PV_RNG_STEP, a.k0, a.k1, sw);
                        if (pv_u32_to_float(sw[1]) > .5f) { Tr = 0.f; stop = true; break; }
                        Tr = Tr * 2.f;                                     // Tr /= continueProb (0.5): exact
                    }
                }
                const float s_dens = __shfl_sync(PV_FULL, c_dens, i);
                Lv = Lv + Tr * (le * s_dens);                              // Lv += Tr * vr->Lve(p, w, time)
                if (SINGLE) {
                    float s_dfac = __shfl_sync(PV_FULL, c_dfac, i);
                    if (s_dfac != 0.f) {                                   // lit, unoccluded, sigma_s(p) != 0
                        const pv_light &lt = sc.lights[__shfl_sync(PV_FULL, c_ln, i)];
                        const float I = bin ? lt.intensity[lane] : 0.f;
                        if (rainbow) {
                            // the march kernel leaves the phase function out for rainbow media (their photon-volume direct term
                            // is rainbowReflection); RainbowVolume::p is HomogeneousVolumeDensity's PhaseHG
                            const v3 sp = ray_at(ro, rd, __shfl_sync(PV_FULL, c_t, i));
                            s_dfac = (s_dfac * phase_hg(-rd, -vi_light_wo(lt, sp), med.g)) * (float)nLights;
                        }
                        const float Ld = (I * expf(-(sig_t * __shfl_sync(PV_FULL, c_sh, i)))) * s_dfac;
                        Lv = Lv + (Tr * (sig_s * s_dens)) * Ld;            // Tr * ss * p * Ld * nLights / pdf
                    }
                }
            }
        }
    }
    if (bin) { a.L[ri * PV_NSPEC + lane] = Lv * step; a.T[ri * PV_NSPEC + lane] = Tr; }
}

#define VT_THREADS 128
// One thread per ray (frames of camera rays): same operations on the same values as volint_kernel, bin loops unrolled over
// register arrays.  Spectra of the medium and of the lights are staged in shared memory (every lane reads the same word).
template <bool SINGLE>
__global__ void __launch_bounds__(VT_THREADS) volint_thread_kernel(VolIntArgs a) {
    __shared__ float s_sig_t[PV_NSPEC], s_sig_s[PV_NSPEC], s_le[PV_NSPEC], s_cy[PV_NSPEC];
    __shared__ float s_I[PV_MAX_LIGHTS][PV_NSPEC];
    const DevScene &sc = *a.sc;
    const DevMedium &med = sc.med;
    const int nLights = (int)sc.n_lights;
    for (int i = threadIdx.x; i < PV_NSPEC; i += VT_THREADS) {
        s_sig_t[i] = med.sigma_a[i] + med.sigma_s[i]; s_sig_s[i] = med.sigma_s[i]; s_le[i] = med.le[i]; s_cy[i] = sc.cie_y[i];
    }
    if (SINGLE)
        for (int i = threadIdx.x; i < nLights * PV_NSPEC; i += VT_THREADS) s_I[i / PV_NSPEC][i % PV_NSPEC] = sc.lights[i / PV_NSPEC].intensity[i % PV_NSPEC];
    __syncthreads();
    const uint64_t ri = (uint64_t)blockIdx.x * VT_THREADS + threadIdx.x;
    if (ri >= a.n) return;
    float sig_t_max = 0.f;
#pragma unroll
    for (int b = 0; b < PV_NSPEC; ++b) sig_t_max = fmaxf(sig_t_max, s_sig_t[b]);
    const bool rainbow = med.type == PV_MEDIUM_RAINBOW;

    const float4 h0 = __ldg(reinterpret_cast<const float4 *>(a.hdr + ri));
    const int nSamples = __float_as_int(h0.z);
    const float step = h0.w;
    const float4 *rp = reinterpret_cast<const float4 *>(a.steps + (((unsigned long long)__float_as_uint(h0.y) << 32) | __float_as_uint(h0.x)));
    float Tr[PV_NSPEC], Lv[PV_NSPEC];
#pragma unroll
    for (int b = 0; b < PV_NSPEC; ++b) { Tr[b] = 1.f; Lv[b] = 0.f; }
    if (nSamples > 0) {
        const v3 ro = V3(__ldg(&a.rays[ri].o[0]), __ldg(&a.rays[ri].o[1]), __ldg(&a.rays[ri].o[2]));
        const v3 rd = V3(__ldg(&a.rays[ri].d[0]), __ldg(&a.rays[ri].d[1]), __ldg(&a.rays[ri].d[2]));
        const uint64_t gidx = a.ray_index ? a.ray_index[ri] : a.ray_index_base + ri;
        float S = 0.f;
        float4 na = __ldg(rp), nb = __ldg(rp + 1);                        // t, tau, rr, dens | sh, dfac, ln, ray
        for (int i = 0; i < nSamples; ++i, rp += 2) {
            const float4 ra = na, rb = nb;
            if (i + 1 < nSamples) { na = __ldg(rp + 2); nb = __ldg(rp + 3); }   // the next record is in flight during this step's bins
            const float s_tau = ra.y, s_dens = ra.w;
#pragma unroll
            for (int b = 0; b < PV_NSPEC; ++b) Tr[b] = Tr[b] * expf(-(s_sig_t[b] * s_tau));
            S += s_tau;
            if (sig_t_max * S > 6.0f) {
                float yy = 0.f;
#pragma unroll
                for (int b = 0; b < PV_NSPEC; ++b) yy += s_cy[b] * Tr[b];
                if (__fdiv_rn(yy * 300.f, 106.856895f * (float)PV_NSPEC) < 1e-3f) {
                    uint32_t sw[4];
                    pv_philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)i, PV_RNG_STEP, a.k0, a.k1, sw);
                    if (pv_u32_to_float(sw[1]) > .5f) {
#pragma unroll
                        for (int b = 0; b < PV_NSPEC; ++b) Tr[b] = 0.f;
                        break;
                    }
#pragma unroll
                    for (int b = 0; b < PV_NSPEC; ++b) Tr[b] = Tr[b] * 2.f;
                }
            }
#pragma unroll
            for (int b = 0; b < PV_NSPEC; ++b) Lv[b] = Lv[b] + Tr[b] * (s_le[b] * s_dens);
            if (SINGLE) {
                float s_dfac = rb.y;
                if (s_dfac != 0.f) {
                    const int ln = __float_as_int(rb.z);
                    if (rainbow) {
                        const v3 sp = ray_at(ro, rd, ra.x);
                        s_dfac = (s_dfac * phase_hg(-rd, -vi_light_wo(sc.lights[ln], sp), med.g)) * (float)nLights;
                    }
                    const float s_sh = rb.x;
                    const float *I = s_I[ln];
#pragma unroll
                    for (int b = 0; b < PV_NSPEC; ++b) {
                        const float Ld = (I[b] * expf(-(s_sig_t[b] * s_sh))) * s_dfac;
                        Lv[b] = Lv[b] + (Tr[b] * (s_sig_s[b] * s_dens)) * Ld;
                    }
                }
            }
        }
    }
    float *Lo = a.L + ri * PV_NSPEC, *To = a.T + ri * PV_NSPEC;
#pragma unroll
    for (int b = 0; b < PV_NSPEC; ++b) { Lo[b] = Lv[b] * step; To[b] = Tr[b]; }
}

static int volint_slice(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, int integrator, float *d_L, float *d_T) {
    uint64_t total = 0;
    pv_gather_params p = *prm;
    const uint32_t flags = PV_GATHER_NO_INDIRECT | (integrator == PV_VOLINT_EMISSION ? PV_GATHER_NO_DIRECT : 0u);
    int rc = pvi_march(ctx, d_rays, n, &p, flags, &total);
    if (rc == PV_ENOMEM && n > 4096) {                      // step records do not fit: two half slices
        const uint64_t h = n / 2;
        rc = volint_slice(ctx, d_rays, h, prm, integrator, d_L, d_T); if (rc) return rc;
        p.ray_index_base = prm->ray_index_base + h;
        return volint_slice(ctx, d_rays + h, n - h, &p, integrator, d_L + h * PV_NSPEC, d_T + h * PV_NSPEC);
    }
    if (rc == PV_ENOMEM) ctx->err = "pv_volume_li: out of device memory for the march records";
    if (rc) return rc;
    VolIntArgs a;
    a.sc = ctx->dscene; a.rays = d_rays; a.hdr = (const RayHdr *)ctx->march_hdr; a.steps = (const StepRec *)ctx->march_steps; a.n = n;
    a.k0 = (uint32_t)prm->seed; a.k1 = (uint32_t)(prm->seed >> 32); a.ray_index_base = prm->ray_index_base;
    a.ray_index = ctx->d_ray_index ? ctx->d_ray_index + (d_rays - ctx->ray_index_rays) : nullptr;
    a.L = d_L; a.T = d_T;
    // Which schedule?  One thread per ray once the rays alone fill the machine (a few resident warps on every SM), one warp per
    // ray below that; PV_VOLINT_THREAD_PER_RAY / PV_VOLINT_WARP_PER_RAY in params->flags force one or the other.
    bool per_thread = n >= (uint64_t)ctx->sm_count * 512;
    if (prm->flags & PV_VOLINT_THREAD_PER_RAY) per_thread = true;
    if (prm->flags & PV_VOLINT_WARP_PER_RAY) per_thread = false;
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (per_thread) {
        const uint32_t blocks = (uint32_t)((n + VT_THREADS - 1) / VT_THREADS);
        if (integrator == PV_VOLINT_EMISSION) volint_thread_kernel<false><<<blocks, VT_THREADS, 0, ctx->stream>>>(a);
        else volint_thread_kernel<true><<<blocks, VT_THREADS, 0, ctx->stream>>>(a);
    } else {
        const uint32_t blocks = (uint32_t)((n * 32 + VI_THREADS - 1) / VI_THREADS);
        if (integrator == PV_VOLINT_EMISSION) volint_kernel<false><<<blocks, VI_THREADS, 0, ctx->stream>>>(a);
        else volint_kernel<true><<<blocks, VI_THREADS, 0, ctx->stream>>>(a);
    }
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->ev1));     // the next slice overwrites the march records
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms += ms;
    cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3); ctx->last_march_ms += ms;
    return PV_OK;
}

// Li of rays [0, n) (device pointers) for VolumeIntegrator "single" / "emission".  Only stepsize, seed and ray_index_base of
// the parameters are read; no photon map is involved.
int pvi_volume_li(pv_ctx *ctx, int integrator, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, float *d_L, float *d_T) {
    if (!ctx->has_scene) { ctx->err = "pv_volume_li: no scene"; return PV_ESTATE; }
    if (integrator != PV_VOLINT_SINGLE && integrator != PV_VOLINT_EMISSION) { ctx->err = "pv_volume_li: unknown integrator"; return PV_EINVAL; }
    if (!(prm->stepsize > 0.f)) { ctx->err = "pv_volume_li: stepsize must be > 0"; return PV_EINVAL; }
    ctx->last_ms = 0.f; ctx->last_march_ms = 0.f;
    const uint64_t slice = PV_GATHER_SLICE_RAYS;
    for (uint64_t r0 = 0; r0 < n; r0 += slice) {
        const uint64_t nr = std::min<uint64_t>(slice, n - r0);
        pv_gather_params p = *prm; p.ray_index_base = prm->ray_index_base + r0;
        int rc = volint_slice(ctx, d_rays + r0, nr, &p, integrator, d_L + r0 * PV_NSPEC, d_T + r0 * PV_NSPEC);
        if (rc) return rc;
    }
    return PV_OK;
}
