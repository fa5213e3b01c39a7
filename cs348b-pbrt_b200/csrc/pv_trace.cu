// pv_trace.cu -- batch forms of Scene::Intersect / IntersectP (accelerators/bvh.cpp:585-685 walking the
// exported LinearBVHNode array, shapes/trianglemesh.cpp:127-281) and of
// PhotonVolumeIntegrator::Transmittance (integrators/photonvolume.cpp:15-30).  One thread per ray.
#include "pv_ctx.h"

template <bool SPH>
__global__ void intersect_kernel(const DevScene *__restrict__ sc, const pv_ray *__restrict__ rays, uint64_t n, uint32_t *__restrict__ prim,
                                 float *__restrict__ t) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pv_ray r = rays[i];
    float maxt = r.maxt;
    int h = bvh_traverse<false, SPH>(*sc, V3(r.o[0], r.o[1], r.o[2]), V3(r.d[0], r.d[1], r.d[2]), r.mint, &maxt, nullptr);
    prim[i] = h < 0 ? 0xFFFFFFFFu : (uint32_t)h;
    t[i] = h < 0 ? INFINITY : maxt;
}
template <bool SPH>
__global__ void occluded_kernel(const DevScene *__restrict__ sc, const pv_ray *__restrict__ rays, uint64_t n, uint8_t *__restrict__ hit) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pv_ray r = rays[i];
    float maxt = r.maxt;
    hit[i] = bvh_traverse<true, SPH>(*sc, V3(r.o[0], r.o[1], r.o[2]), V3(r.d[0], r.d[1], r.d[2]), r.mint, &maxt, nullptr) >= 0 ? 1 : 0;
}
// one thread per ray: the optical-depth scalar is marched once, then the 30 bins of exp(-sigma_t s) go out (neighbouring threads
// are neighbouring rays: their density taps share sectors)
__global__ void transmittance_kernel(const DevScene *__restrict__ sc, const pv_ray *__restrict__ rays, uint64_t n, float step,
                                     const float *__restrict__ u, float *__restrict__ T) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevMedium &med = sc->med;
    float s = 0.f;
    const bool has = med.type != PV_MEDIUM_NONE;
    if (has) {
        const pv_ray r = rays[i];
        s = med_tau_scalar(med, V3(r.o[0], r.o[1], r.o[2]), V3(r.d[0], r.d[1], r.d[2]), r.mint, r.maxt, step, u ? u[i] : 0.5f, nullptr);
    }
    float *out = T + i * PV_NSPEC;
#pragma unroll 1
    for (int b = 0; b < PV_NSPEC; ++b) out[b] = has ? expf(-((med.sigma_a[b] + med.sigma_s[b]) * s)) : 1.f;
}

int pvi_intersect(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, uint32_t *d_prim, float *d_t) {
    if (!ctx->has_scene) { ctx->err = "pv_intersect: no scene"; return PV_ESTATE; }
    if (!n) return PV_OK;
    if (ctx->hscene.n_spheres) intersect_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->dscene, d_rays, n, d_prim, d_t);
    else intersect_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->dscene, d_rays, n, d_prim, d_t);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
int pvi_occluded(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, uint8_t *d_hit) {
    if (!ctx->has_scene) { ctx->err = "pv_occluded: no scene"; return PV_ESTATE; }
    if (!n) return PV_OK;
    if (ctx->hscene.n_spheres) occluded_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->dscene, d_rays, n, d_hit);
    else occluded_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->dscene, d_rays, n, d_hit);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
int pvi_transmittance(pv_ctx *ctx, const pv_ray *d_rays, uint64_t n, float step, const float *d_u, float *d_T) {
    if (!ctx->has_scene) { ctx->err = "pv_transmittance: no scene"; return PV_ESTATE; }
    if (!n) return PV_OK;
    transmittance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->dscene, d_rays, n, step, d_u, d_T);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    return PV_OK;
}
