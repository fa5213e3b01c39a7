// pv_march.cu -- the map-independent half of PhotonVolumeIntegrator::Li (integrators/photonvolume.cpp:112-222).
//
//   march_setup_kernel  one thread per camera ray: medium interval (vr->IntersectP :119-124), nSamples / step
//                       (:130-131), jittered start (:135); reserves the ray's StepRec run with one atomicAdd per warp.
//   march_steps_kernel  one thread per ray (a warp = 32 neighbouring rays marching in lockstep): sample position
//                       with the reference's accumulated t0 += step (:147), step-segment optical depth (:153-155),
//                       Russian-roulette draw (:158-165), density at the sample, light choice (:177-182), shadow ray
//                       through the LinearBVHNode array and its optical depth (VisibilityTester::Transmittance,
//                       core/light.cpp:51-56), phase function.
// Both are streaming kernels: no shared memory, full occupancy.  The photon lookups and the Lv/Tr recurrence are in
// pv_gather.cu; seven scalars per step cross over in a StepRec (pv_march.cuh).
#include <algorithm>
#include "pv_ctx.h"
#include "pv_march.cuh"

#ifndef MS_THREADS
#define MS_THREADS 128
#endif
#ifndef MS_MIN_CTAS
#define MS_MIN_CTAS 8
#endif

__global__ void __launch_bounds__(MS_THREADS) march_setup_kernel(const DevScene *__restrict__ scp, const pv_ray *__restrict__ rays, uint64_t n,
                                                                float stepsize, RayHdr *__restrict__ hdr, unsigned long long *total,
                                                                uint32_t *__restrict__ block_cost) {
    const DevMedium &med = scp->med;
    const uint64_t ri = (uint64_t)blockIdx.x * MS_THREADS + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    int nSamples = 0; float step = 0.f, t_first = 0.f, tbase = 0.f;
    if (ri < n) {
        const pv_ray ray = rays[ri];
        const v3 ro = V3(ray.o[0], ray.o[1], ray.o[2]), rd = V3(ray.d[0], ray.d[1], ray.d[2]);
        float t0, t1;
        if (med.type != PV_MEDIUM_NONE && med_intersectp(med, ro, rd, ray.mint, ray.maxt, &t0, &t1) && (t1 - t0) != 0.f) {
            nSamples = (int)ceilf(__fdiv_rn(t1 - t0, stepsize));
            if (nSamples < 0) nSamples = 0;
            step = __fdiv_rn(t1 - t0, (float)nSamples);
            t_first = t0;                                   // p = ray(t0) before the jitter (photonvolume.cpp:133)
            tbase = t0 + ray.u_scatter * step;              // t0 += u * step
        }
    }
    // reserve nSamples records: exclusive scan inside the warp, one atomic per warp
    uint32_t inc = (uint32_t)nSamples;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(PV_FULL, inc, o); if (lane >= o) inc += t; }
    unsigned long long base = 0;
    if (lane == 31 && inc) { base = atomicAdd(total, (unsigned long long)inc); atomicAdd(block_cost + blockIdx.x, inc); }    // (block_cost is zeroed by the host)
    base = __shfl_sync(PV_FULL, base, 31);
    if (ri < n) {
        RayHdr h;
        h.offset = base + (inc - (uint32_t)nSamples); h.nSamples = nSamples; h.step = step;
        h.pad[0] = t_first; h.pad[1] = tbase; h.pad[2] = 0.f; h.pad[3] = 0.f;
        reinterpret_cast<float4 *>(hdr + ri)[0] = reinterpret_cast<const float4 *>(&h)[0];
        reinterpret_cast<float4 *>(hdr + ri)[1] = reinterpret_cast<const float4 *>(&h)[1];
    }
}

// Blocks of MS_THREADS rays in order of decreasing march steps, for the thread-per-ray kernels that follow (march_steps_kernel,
// recurrence_thread_kernel): the hardware hands CTAs out in index order, so with the heaviest blocks first the tail of a launch is
// made of light blocks.  On a full frame (16 k blocks) the tail does not matter; on one GPU's share of a frame spread over eight
// (2 k blocks = 1.7 waves) it was a third of the march.  Counting sort over 64 cost classes, one CTA.
__global__ void __launch_bounds__(1024) block_order_kernel(const uint32_t *__restrict__ cost, uint32_t nblocks, uint32_t *__restrict__ order) {
    __shared__ uint32_t s_max, s_cnt[64], s_off[64];
    if (threadIdx.x == 0) s_max = 1u;
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t mx = 0;
    for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) mx = max(mx, cost[b]);
    atomicMax(&s_max, mx);
    __syncthreads();
    const float scale = 63.999f / (float)s_max;
    for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) atomicAdd(&s_cnt[63u - min(63u, (uint32_t)((float)cost[b] * scale))], 1u);     // class 0 = heaviest
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t acc = 0; for (int c = 0; c < 64; ++c) { s_off[c] = acc; acc += s_cnt[c]; } }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) order[atomicAdd(&s_off[63u - min(63u, (uint32_t)((float)cost[b] * scale))], 1u)] = b;
}

__global__ void publish_total_kernel(const unsigned long long *total, unsigned long long *host_mapped) { *host_mapped = *total; }

struct MarchArgs {
    const DevScene *sc;
    const pv_ray *rays;
    const RayHdr *hdr;
    StepRec *steps;
    uint64_t n;
    float stepsize;
    uint32_t flags;
    uint32_t k0, k1;               // Philox key
    uint64_t ray_index_base;
    const uint64_t *ray_index;     // per-ray stream indices (pv_gather_indexed) or null: ray i draws from ray_index_base + i
    const uint32_t *block_order;   // CTA i takes the rays of block block_order[i] (block_order_kernel)
    pv_gather_stats *stats;
};

// One THREAD per ray, steps in sequence: the 32 lanes of a warp are 32 neighbouring camera rays (callers pass rays
// in image-tile order) at the same march depth, so their density taps fall into the same few voxels and the warp's
// eight trilinear taps touch a handful of 32-byte sectors instead of 256.
template <bool SPH>
__global__ void __launch_bounds__(MS_THREADS, MS_MIN_CTAS) march_steps_kernel(MarchArgs a) {
    const DevScene &sc = *a.sc;
    const DevMedium &gmed = sc.med;
    const MedView med = make_medview<MedView>(gmed);             // extent / grid dimensions / grid pointer in registers
    const uint32_t lane = threadIdx.x & 31;
    float sig_t_max = 0.f; bool any_sig_s = false;
    for (int bb = 0; bb < PV_NSPEC; ++bb) { sig_t_max = fmaxf(sig_t_max, gmed.sigma_a[bb] + gmed.sigma_s[bb]); any_sig_s |= gmed.sigma_s[bb] != 0.f; }
    const bool rainbow = med.type == PV_MEDIUM_RAINBOW;
    const bool do_direct = any_sig_s && sc.n_lights > 0 && !(a.flags & PV_GATHER_NO_DIRECT);
    const int nLights = (int)sc.n_lights;
    uint32_t ns = 0, nshadow = 0;
    const uint64_t ri = (uint64_t)a.block_order[blockIdx.x] * MS_THREADS + threadIdx.x;
    int nSamples = 0;
    float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
    if (ri < a.n) {
        h0 = __ldg(reinterpret_cast<const float4 *>(a.hdr + ri)); h1 = __ldg(reinterpret_cast<const float4 *>(a.hdr + ri) + 1);
        nSamples = __float_as_int(h0.z);
    }
    if (nSamples > 0) {
        const unsigned long long off = ((unsigned long long)__float_as_uint(h0.y) << 32) | __float_as_uint(h0.x);
        const float step = h0.w, t_first = h1.x;
        const pv_ray ray = a.rays[ri];
        const v3 ro = V3(ray.o[0], ray.o[1], ray.o[2]), rd = V3(ray.d[0], ray.d[1], ray.d[2]);
        const uint64_t gidx = a.ray_index ? a.ray_index[ri] : a.ray_index_base + ri;
        uint32_t rw[4] = {0u, 0u, 0u, 0u};
        if (do_direct) pv_philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, PV_RNG_RAY, a.k0, a.k1, rw);
        float c_t = h1.y;                                   // t0 + u * step; then t0 += step per sample (photonvolume.cpp:135,147)
        v3 pPrev = ray_at(ro, rd, t_first);
        float4 *out = reinterpret_cast<float4 *>(a.steps + off);
        for (int si = 0; si < nSamples; ++si, c_t += step, out += 2) {
            float c_tau, c_rr = -1.f, c_dens, c_sh = 0.f, c_dfac = 0.f;
            int c_ln = 0;
            const v3 p = ray_at(ro, rd, c_t);
            uint32_t sw[4];
            pv_philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)si, PV_RNG_STEP, a.k0, a.k1, sw);
            c_tau = med_tau_scalar(med, pPrev, p - pPrev, 0.f, 1.f, .5f * a.stepsize, pv_u32_to_float(sw[0]), &ns);
            pPrev = p;
            // Tr.y() < 1e-3 ?  exp(-sig_t_max * tau) bounds every bin from below and y(1) ~ 1, so only large taus need the sum
            if (sig_t_max * c_tau > 6.0f) {
                float yy = 0.f;
                for (int bb = 0; bb < PV_NSPEC; ++bb) yy += sc.cie_y[bb] * expf(-((gmed.sigma_a[bb] + gmed.sigma_s[bb]) * c_tau));
                if (__fdiv_rn(yy * 300.f, 106.856895f * (float)PV_NSPEC) < 1e-3f) c_rr = pv_u32_to_float(sw[1]);
            }
            c_dens = med_density(med, p, &ns);
            if (c_dens != 0.f && do_direct) {
                const float u_l = pv_van_der_corput(pv_permute((uint32_t)si, (uint32_t)nSamples, rw[1]), rw[0]);
                c_ln = min((int)floorf(u_l * nLights), nLights - 1);
                if (sc.lights[c_ln].type == PV_LIGHT_AREA) {
                    // DiffuseAreaLight: the light sample of the step is (lightComp, lightPos[0], lightPos[1]) handed to LightSample in
                    // the integrators' argument order -- LightSample(up0, up1, ucomp) called as ls(comp, pos0, pos1), i.e. the
                    // component number is the THIRD number (integrators/single.cpp:116, photonvolume.cpp:183); three words of a
                    // Philox block of its own here
                    uint32_t aw[4];
                    pv_philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), (uint32_t)si, PV_RNG_AREA, a.k0, a.k1, aw);
                    v3 wi, vis_d; float pdf, vis_maxt;
                    const bool facing = area_sample_L(sc, c_ln, p, pv_u32_to_float(aw[2]), pv_u32_to_float(aw[0]), pv_u32_to_float(aw[1]), &wi, &pdf,
                                                      &vis_d, &vis_maxt);
                    if (facing && pdf > 0.f) {
                        nshadow++;
                        float mt = vis_maxt;
                        if (bvh_traverse<true, SPH>(sc, p, vis_d, 0.f, &mt, nullptr) < 0) {
                            c_sh = med_tau_scalar(med, p, vis_d, 0.f, vis_maxt, 4.f * a.stepsize, pv_u32_to_float(sw[2]), &ns);
                            c_dfac = __fdiv_rn(med_phase(med, p, -rd, -wi) * (float)nLights, pdf);      // p * Ld * nLights / pdf, Ld = Lemit * Tr
                        }
                    }
                } else {
                    LightQuery lq;
                    light_query(sc.lights[c_ln], p, &lq);
                    if (lq.falloff != 0.f) {
                        nshadow++;
                        float mt = lq.vis_maxt;
                        if (bvh_traverse<true, SPH>(sc, lq.vis_o, lq.vis_d, lq.vis_mint, &mt, nullptr) < 0) {
                            c_sh = med_tau_scalar(med, lq.vis_o, lq.vis_d, lq.vis_mint, lq.vis_maxt, 4.f * a.stepsize, pv_u32_to_float(sw[2]), &ns);
                            const float geom = lq.point_like ? __fdiv_rn(lq.falloff, lq.inv_mode_d2) : 1.f;
                            c_dfac = rainbow ? geom : (geom * med_phase(med, p, -rd, -lq.wi)) * (float)nLights;
                        }
                    }
                }
            }
            out[0] = make_float4(c_t, c_tau, c_rr, c_dens);
            out[1] = make_float4(c_sh, c_dfac, __int_as_float(c_ln), __uint_as_float((uint32_t)ri));    // pad = the ray of the step (slice-relative)
            if (c_rr > .5f && !(a.flags & PV_GATHER_NO_INDIRECT)) {   // the roulette ends the photon-volume march here (the volint kernels
                                                                     // decide on the cumulative Tr instead): the remaining records are dead
                for (++si, out += 2; si < nSamples; ++si, out += 2) {
                    out[0] = make_float4(0.f, 0.f, PV_RR_DEAD, 0.f);
                    out[1] = make_float4(0.f, 0.f, 0.f, __uint_as_float((uint32_t)ri));
                }
                break;
            }
        }
    }
    ns = __reduce_add_sync(PV_FULL, ns); nshadow = __reduce_add_sync(PV_FULL, nshadow);
    if (lane == 0 && a.stats && (ns | nshadow)) {
        atomicAdd((unsigned long long *)&a.stats->density_samples, (unsigned long long)ns);
        atomicAdd((unsigned long long *)&a.stats->shadow_rays, (unsigned long long)nshadow);
    }
}

// Fills ctx->march_hdr / ctx->march_steps for rays [0, n).  Returns PV_ENOMEM (without an error message change) when the
// step records do not fit, so the caller can retry with fewer rays.
int pvi_march(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, const pv_gather_params *prm, uint32_t flags, uint64_t *total_steps) {
    int rc = pv_ensure(ctx, &ctx->march_hdr, &ctx->march_hdr_bytes, n * sizeof(RayHdr)); if (rc) return rc;
    unsigned long long *d_total = ctx->d_counters + 8;
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), ctx->stream));
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
    const uint32_t blocks = (uint32_t)((n + MS_THREADS - 1) / MS_THREADS);
    rc = pv_ensure(ctx, &ctx->march_blk, &ctx->march_blk_bytes, (size_t)blocks * 2 * sizeof(uint32_t)); if (rc) return rc;
    uint32_t *blk_cost = (uint32_t *)ctx->march_blk, *blk_order = blk_cost + blocks;
    PV_CUDA_CHECK(ctx, cudaMemsetAsync(blk_cost, 0, (size_t)blocks * sizeof(uint32_t), ctx->stream));
    march_setup_kernel<<<blocks, MS_THREADS, 0, ctx->stream>>>(ctx->dscene, d_rays, n, prm->stepsize, (RayHdr *)ctx->march_hdr, d_total, blk_cost);
    block_order_kernel<<<1, 1024, 0, ctx->stream>>>(blk_cost, blocks, blk_order);
    ctx->launches += 3;                                 // + publish_total_kernel below
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    // The step count comes back through mapped pinned memory, not a memcpy: a copy would queue on the device->host copy
    // engine behind the result download of the previous slice (pv_gather overlaps the two).
    if (!ctx->h_total) PV_CUDA_CHECK(ctx, cudaHostAlloc((void **)&ctx->h_total, sizeof(unsigned long long), cudaHostAllocMapped));
    publish_total_kernel<<<1, 1, 0, ctx->stream>>>(d_total, ctx->h_total);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned long long total = *(volatile unsigned long long *)ctx->h_total;
    *total_steps = total;
    size_t free_b = 0, total_b = 0;
    const size_t need = (size_t)total * sizeof(StepRec);
    if (need > ctx->march_steps_bytes) {
        PV_CUDA_CHECK(ctx, cudaMemGetInfo(&free_b, &total_b));
        if (need > free_b + ctx->march_steps_bytes || need > (size_t)PV_MARCH_MAX_BYTES) return PV_ENOMEM;
    }
    rc = pv_ensure(ctx, &ctx->march_steps, &ctx->march_steps_bytes, need); if (rc) return rc;
    MarchArgs a;
    a.sc = ctx->dscene; a.rays = d_rays; a.hdr = (const RayHdr *)ctx->march_hdr; a.steps = (StepRec *)ctx->march_steps; a.n = n;
    a.stepsize = prm->stepsize; a.flags = flags; a.k0 = (uint32_t)prm->seed; a.k1 = (uint32_t)(prm->seed >> 32);
    a.ray_index_base = prm->ray_index_base; a.stats = ctx->d_stats;
    a.ray_index = ctx->d_ray_index ? ctx->d_ray_index + (d_rays - ctx->ray_index_rays) : nullptr;
    a.block_order = blk_order;
    if (total) {
        if (ctx->hscene.n_spheres) march_steps_kernel<true><<<blocks, MS_THREADS, 0, ctx->stream>>>(a);
        else march_steps_kernel<false><<<blocks, MS_THREADS, 0, ctx->stream>>>(a);
        ctx->launches += 1;
        PV_CUDA_CHECK(ctx, cudaGetLastError());
    }
    PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
    return PV_OK;
}
