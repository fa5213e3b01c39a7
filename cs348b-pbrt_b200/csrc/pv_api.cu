// pv_api.cu -- the extern "C" layer of include/pv.h: argument checks, host<->device staging,
// context lifetime.  No arithmetic of the path lives here.
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "pv_ctx.h"

static thread_local std::string g_err;

int pv_ensure(pv_ctx *ctx, void **p, size_t *cap, size_t bytes) {
    if (*cap >= bytes && *p) return PV_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t want = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e); *p = nullptr; return PV_ENOMEM; }
    *cap = want;
    return PV_OK;
}

#define LOCK(ctx) if (!(ctx)) { g_err = "null context"; return PV_EINVAL; } std::lock_guard<std::mutex> lock__((ctx)->mu); \
    { cudaError_t e__ = cudaSetDevice((ctx)->device); if (e__ != cudaSuccess) { (ctx)->err = cudaGetErrorString(e__); return PV_ECUDA; } }

extern "C" {

int pv_version(void) { return 100; }

const char *pv_last_error(pv_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int pv_create(pv_ctx **out, int device) {
    if (!out) { g_err = "pv_create: null out"; return PV_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_err = std::string("pv_create: no CUDA device (") + cudaGetErrorString(e) + "); this path has no CPU fallback";
        return PV_ECUDA;
    }
    if (device < 0 || device >= ndev) { g_err = "pv_create: bad device index"; return PV_EINVAL; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return PV_ECUDA; }
    pv_ctx *ctx = new pv_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
              cudaEventCreate(&ctx->ev2) == cudaSuccess && cudaEventCreate(&ctx->ev3) == cudaSuccess &&
              cudaEventCreate(&ctx->tev[0]) == cudaSuccess && cudaEventCreate(&ctx->tev[1]) == cudaSuccess &&
              cudaEventCreate(&ctx->tev[2]) == cudaSuccess && cudaEventCreate(&ctx->tev[3]) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dscene, sizeof(DevScene)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->d_stats, sizeof(pv_gather_stats)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->d_counters, 64 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemset(ctx->d_stats, 0, sizeof(pv_gather_stats)) == cudaSuccess;
    if (!ok) { g_err = std::string("pv_create: ") + cudaGetErrorString(cudaGetLastError()); pv_destroy(ctx); return PV_ECUDA; }   // frees what was made
    *out = ctx;
    return PV_OK;
}

void pv_destroy(pv_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pvi_comm_destroy(ctx);
    void *ptrs[] = {ctx->dscene, ctx->d_nodes, ctx->d_tri, ctx->d_prim_mat, ctx->d_mats, ctx->d_lights, ctx->d_density, ctx->d_spheres, ctx->d_pos, ctx->d_wi,
                    ctx->d_alpha, ctx->d_ids, ctx->m_pos4, ctx->m_wi4, ctx->m_alpha32, ctx->m_orig, ctx->cell_start, ctx->scratch, ctx->io, ctx->io2,
                    ctx->d_stats, ctx->d_counters, ctx->march_hdr, ctx->march_steps, ctx->lii, ctx->cg_sort, ctx->cg_overflow, ctx->sort_hist, ctx->wf, ctx->d_mat_flags, ctx->d_ltris, ctx->d_ltri_area, ctx->d_ltri_cdf, ctx->io3, ctx->march_blk};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int c = 0; c < 4; ++c) pvi_free_set(&ctx->surf[c]);
    if (ctx->rad_Lo) cudaFree(ctx->rad_Lo);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    for (int i = 0; i < 4; ++i) if (ctx->tev[i]) cudaEventDestroy(ctx->tev[i]);
    if (ctx->h_total) cudaFreeHost(ctx->h_total);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void *pv_stream(pv_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

static int upload(pv_ctx *ctx, void **dst, const void *src, size_t bytes) {
    if (*dst) { cudaFree(*dst); *dst = nullptr; }
    if (!bytes) return PV_OK;
    PV_CUDA_CHECK(ctx, cudaMalloc(dst, bytes));
    PV_CUDA_CHECK(ctx, cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return PV_OK;
}

int pv_set_scene(pv_ctx *ctx, const pv_scene_desc *s) {
    LOCK(ctx);
    if (!s) { ctx->err = "pv_set_scene: null scene"; return PV_EINVAL; }
    if (s->n_lights > PV_MAX_LIGHTS) { ctx->err = "pv_set_scene: more than 16 lights"; return PV_EINVAL; }
    if (s->n_prims && (!s->tri_verts || !s->prim_material || !s->materials)) { ctx->err = "pv_set_scene: null geometry arrays"; return PV_EINVAL; }
    for (uint32_t i = 0; i < s->n_prims; ++i)
        if (s->prim_material[i] >= s->n_materials) { ctx->err = "pv_set_scene: material index out of range"; return PV_EINVAL; }
    if (s->n_spheres && (!s->spheres || !s->prim_shape)) { ctx->err = "pv_set_scene: spheres without a prim_shape table"; return PV_EINVAL; }
    for (uint32_t i = 0; s->n_spheres && i < s->n_prims; ++i)
        if (s->prim_shape[i] != PV_SHAPE_TRIANGLE && s->prim_shape[i] >= s->n_spheres) { ctx->err = "pv_set_scene: sphere index out of range"; return PV_EINVAL; }
    for (uint32_t i = 0; i < s->n_lights; ++i) {
        const pv_light &l = s->lights[i];
        if (l.type != PV_LIGHT_AREA) continue;
        if (!s->light_tris || (uint64_t)l.area.first_tri + l.area.n_tris > s->n_light_tris || l.area.n_tris == 0) {
            ctx->err = "pv_set_scene: area light without triangles in light_tris"; return PV_EINVAL;
        }
        if (s->medium && s->medium->type == PV_MEDIUM_RAINBOW) { ctx->err = "pv_set_scene: area lights in a rainbow medium are not on this path"; return PV_EINVAL; }
    }
    // the flattened BVH (accelerators/bvh.cpp:154-164): every kernel walks it with a fixed 64-entry todo stack (bvh_traverse,
    // like the reference's todo[64]) and indexes by child / primitive offsets, so a malformed or too deep tree is refused here
    if (s->n_nodes) {
        if (!s->nodes) { ctx->err = "pv_set_scene: null BVH nodes"; return PV_EINVAL; }
        std::vector<std::pair<uint32_t, uint32_t>> todo(1, std::make_pair(0u, 0u));       // node, entries on the traversal stack when it is visited
        std::vector<uint8_t> seen(s->n_nodes, 0);
        while (!todo.empty()) {
            const uint32_t i = todo.back().first, depth = todo.back().second; todo.pop_back();
            if (i >= s->n_nodes || seen[i]) { ctx->err = "pv_set_scene: BVH child offset out of range or shared"; return PV_EINVAL; }
            seen[i] = 1;
            const pv_bvh_node &nd = s->nodes[i];
            if (nd.n_primitives) {
                if ((uint64_t)nd.offset + nd.n_primitives > s->n_prims) { ctx->err = "pv_set_scene: BVH leaf primitive range out of bounds"; return PV_EINVAL; }
            } else {
                if (nd.axis > 2) { ctx->err = "pv_set_scene: BVH split axis > 2"; return PV_EINVAL; }
                if (nd.offset <= i + 1 || nd.offset >= s->n_nodes) { ctx->err = "pv_set_scene: BVH second child offset out of range"; return PV_EINVAL; }
                if (depth + 1 > 64) { ctx->err = "pv_set_scene: BVH deeper than the 64-entry traversal stack"; return PV_EINVAL; }
                todo.push_back(std::make_pair(i + 1, depth + 1)); todo.push_back(std::make_pair(nd.offset, depth + 1));
            }
        }
    }
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_scene = false; ctx->built = false; ctx->shoot_yield[0] = ctx->shoot_yield[1] = 0.;
    int rc;
    if ((rc = upload(ctx, &ctx->d_nodes, s->nodes, sizeof(pv_bvh_node) * (size_t)s->n_nodes))) return rc;
    if (s->n_spheres) {
        // sphere primitives are tagged in the device copy of the triangle table (see DevScene)
        if (s->n_spheres > PV_SPHERE_INDEX_MASK) { ctx->err = "pv_set_scene: too many spheres"; return PV_EINVAL; }
        std::vector<float> tri(s->tri_verts, s->tri_verts + 9 * (size_t)s->n_prims);
        for (uint32_t i = 0; i < s->n_prims; ++i) {
            if (s->prim_shape[i] == PV_SHAPE_TRIANGLE) {
                if (tri[9 * (size_t)i] != tri[9 * (size_t)i]) { ctx->err = "pv_set_scene: NaN triangle vertex"; return PV_EINVAL; }
                continue;
            }
            const uint32_t bits = PV_SPHERE_TAG | s->prim_shape[i];
            memcpy(&tri[9 * (size_t)i], &bits, sizeof(bits));
        }
        if ((rc = upload(ctx, &ctx->d_tri, tri.data(), sizeof(float) * tri.size()))) return rc;
    } else if ((rc = upload(ctx, &ctx->d_tri, s->tri_verts, sizeof(float) * 9 * (size_t)s->n_prims))) return rc;
    if ((rc = upload(ctx, &ctx->d_prim_mat, s->prim_material, sizeof(uint32_t) * (size_t)s->n_prims))) return rc;
    if ((rc = upload(ctx, &ctx->d_mats, s->materials, sizeof(pv_material) * (size_t)s->n_materials))) return rc;
    {
        std::vector<uint8_t> mf(std::max<uint32_t>(s->n_materials, 1u), 0);
        for (uint32_t i = 0; i < s->n_materials; ++i)
            for (int b = 0; b < PV_NSPEC; ++b)
                mf[i] |= (s->materials[i].kd[b] != 0.f ? PV_MATF_KD : 0) | (s->materials[i].kr[b] != 0.f ? PV_MATF_KR : 0) | (s->materials[i].kt[b] != 0.f ? PV_MATF_KT : 0);
        if ((rc = upload(ctx, &ctx->d_mat_flags, mf.data(), mf.size()))) return rc;
    }
    if ((rc = upload(ctx, &ctx->d_lights, s->lights, sizeof(pv_light) * (size_t)s->n_lights))) return rc;
    if ((rc = upload(ctx, &ctx->d_spheres, s->spheres, sizeof(pv_sphere) * (size_t)s->n_spheres))) return rc;
    // area lights: triangle areas (shapes/trianglemesh.cpp:284-290), ShapeSet's area distribution (core/light.cpp:129-136,
    // Distribution1D core/montecarlo.h:55-83) and the sums ShapeSet::Pdf uses, in the reference's float operations
    std::vector<float> lt_area(s->n_light_tris, 0.f), lt_cdf;
    float la_sum[PV_MAX_LIGHTS] = {0}, la_pd[PV_MAX_LIGHTS] = {0}; uint32_t la_off[PV_MAX_LIGHTS] = {0};
    for (uint32_t i = 0; i < s->n_lights; ++i) {
        const pv_light &l = s->lights[i];
        if (l.type != PV_LIGHT_AREA) continue;
        const uint32_t n = l.area.n_tris;
        float sumArea = 0.f, pd = 0.f;
        for (uint32_t k = 0; k < n; ++k) {
            const float *tv = s->light_tris + 9 * (size_t)(l.area.first_tri + k);
            const double ax = (double)(tv[3] - tv[0]), ay = (double)(tv[4] - tv[1]), az = (double)(tv[5] - tv[2]);
            const double bx = (double)(tv[6] - tv[0]), by = (double)(tv[7] - tv[1]), bz = (double)(tv[8] - tv[2]);
            const float cx = (float)(ay * bz - az * by), cy = (float)(az * bx - ax * bz), cz = (float)(ax * by - ay * bx);      // Cross() in double, geometry.h:477-484
            const float ar = 0.5f * sqrtf(cx * cx + cy * cy + cz * cz);
            lt_area[l.area.first_tri + k] = ar; sumArea += ar;
        }
        la_off[i] = (uint32_t)lt_cdf.size();
        lt_cdf.resize(lt_cdf.size() + n + 1);
        float *cdf = lt_cdf.data() + la_off[i];
        cdf[0] = 0.f;
        for (uint32_t k = 1; k < n + 1; ++k) cdf[k] = cdf[k - 1] + lt_area[l.area.first_tri + k - 1] / n;
        const float funcInt = cdf[n];
        if (funcInt == 0.f) for (uint32_t k = 1; k < n + 1; ++k) cdf[k] = (float)k / (float)n;
        else for (uint32_t k = 1; k < n + 1; ++k) cdf[k] /= funcInt;
        for (uint32_t k = 0; k < n; ++k) pd += lt_area[l.area.first_tri + k] * (1.f / lt_area[l.area.first_tri + k]);
        la_sum[i] = sumArea; la_pd[i] = pd;
    }
    if ((rc = upload(ctx, &ctx->d_ltris, s->light_tris, sizeof(float) * 9 * (size_t)s->n_light_tris))) return rc;
    if ((rc = upload(ctx, &ctx->d_ltri_area, lt_area.data(), sizeof(float) * lt_area.size()))) return rc;
    if ((rc = upload(ctx, &ctx->d_ltri_cdf, lt_cdf.data(), sizeof(float) * lt_cdf.size()))) return rc;
    DevScene &h = ctx->hscene;
    memset(&h, 0, sizeof(h));
    h.ltris = (const float *)ctx->d_ltris; h.ltri_area = (const float *)ctx->d_ltri_area; h.ltri_cdf = (const float *)ctx->d_ltri_cdf;
    memcpy(h.larea_sum, la_sum, sizeof(la_sum)); memcpy(h.larea_pd, la_pd, sizeof(la_pd)); memcpy(h.lcdf_off, la_off, sizeof(la_off));
    h.nodes = (const pv_bvh_node *)ctx->d_nodes; h.n_nodes = s->n_nodes;
    h.tri = (const float *)ctx->d_tri; h.prim_mat = (const uint32_t *)ctx->d_prim_mat; h.n_prims = s->n_prims;
    h.mats = (const pv_material *)ctx->d_mats; h.n_mats = s->n_materials;
    h.lights = (const pv_light *)ctx->d_lights; h.n_lights = s->n_lights;
    h.spheres = (const pv_sphere *)ctx->d_spheres; h.n_spheres = s->n_spheres;
    h.mat_flags = (const uint8_t *)ctx->d_mat_flags;
    memcpy(h.world_bound, s->world_bound, sizeof(h.world_bound));
    memcpy(h.cie_y, s->cie_y, sizeof(h.cie_y));
    if (s->medium && s->medium->type != PV_MEDIUM_NONE) {
        const pv_medium *m = s->medium;
        DevMedium &d = h.med;
        d.type = m->type;
        memcpy(d.w2v, m->world_to_volume, sizeof(d.w2v));
        memcpy(d.p0, m->p0, sizeof(d.p0)); memcpy(d.p1, m->p1, sizeof(d.p1));
        memcpy(d.sigma_a, m->sigma_a, sizeof(d.sigma_a)); memcpy(d.sigma_s, m->sigma_s, sizeof(d.sigma_s)); memcpy(d.le, m->le, sizeof(d.le));
        d.g = m->g; d.nx = m->nx; d.ny = m->ny; d.nz = m->nz;
        static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        d.identity = memcmp(d.w2v, ident, sizeof(ident)) == 0;
        if (m->type == PV_MEDIUM_EXPONENTIAL && (!m->density || m->nx != 5 || m->ny != 1 || m->nz != 1)) {
            ctx->err = "pv_set_scene: exponential medium wants density = {a, b, updir.xyz}, nx = 5, ny = nz = 1"; return PV_EINVAL;
        }
        if (m->type == PV_MEDIUM_GRID || m->type == PV_MEDIUM_EXPONENTIAL) {
            if (!m->density || m->nx < 1 || m->ny < 1 || m->nz < 1) { ctx->err = "pv_set_scene: grid medium without density"; return PV_EINVAL; }
            if ((uint64_t)m->nx * (uint64_t)m->ny * (uint64_t)m->nz >= (1ull << 31)) { ctx->err = "pv_set_scene: density grid of 2^31 voxels or more"; return PV_EINVAL; }
            if ((rc = upload(ctx, &ctx->d_density, m->density, sizeof(float) * (size_t)m->nx * m->ny * m->nz))) return rc;
            d.density = (const float *)ctx->d_density;
        }
    } else h.med.type = PV_MEDIUM_NONE;
    // light power CDF: ComputeLightSamplingCDF (core/integrator.cpp:261-268) + Distribution1D (core/montecarlo.h:55-83)
    {
        int n = (int)s->n_lights;
        h.light_cdf[0] = 0.f;
        for (int i = 0; i < n; ++i) h.light_func[i] = s->lights[i].power_y;
        for (int i = 1; i < n + 1; ++i) h.light_cdf[i] = h.light_cdf[i - 1] + h.light_func[i - 1] / n;
        h.light_func_int = n ? h.light_cdf[n] : 0.f;
        if (n) {
            if (h.light_func_int == 0.f) for (int i = 1; i < n + 1; ++i) h.light_cdf[i] = float(i) / float(n);
            else for (int i = 1; i < n + 1; ++i) h.light_cdf[i] /= h.light_func_int;
        }
    }
    PV_CUDA_CHECK(ctx, cudaMemcpy(ctx->dscene, &h, sizeof(h), cudaMemcpyHostToDevice));
    ctx->has_scene = true;
    return PV_OK;
}

static int reserve_photons(pv_ctx *ctx, uint64_t n) {
    if (n <= ctx->cap_photons) return PV_OK;
    uint64_t cap = std::max<uint64_t>(n, 1024);
    float *np = nullptr, *nw = nullptr, *na = nullptr; uint64_t *ni = nullptr;
    if (cudaMalloc((void **)&np, cap * 3 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&nw, cap * 3 * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&na, cap * 32 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&ni, cap * sizeof(uint64_t)) != cudaSuccess) {
        cudaGetLastError();
        if (np) cudaFree(np); if (nw) cudaFree(nw); if (na) cudaFree(na); if (ni) cudaFree(ni);      // nothing of a failed group stays behind
        ctx->err = "out of device memory for " + std::to_string(cap) + " photons"; return PV_ENOMEM;
    }
    if (ctx->n_photons) {
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(np, ctx->d_pos, ctx->n_photons * 3 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(nw, ctx->d_wi, ctx->n_photons * 3 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(na, ctx->d_alpha, ctx->n_photons * 32 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ni, ctx->d_ids, ctx->n_photons * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->d_pos) cudaFree(ctx->d_pos);
    if (ctx->d_wi) cudaFree(ctx->d_wi);
    if (ctx->d_alpha) cudaFree(ctx->d_alpha);
    if (ctx->d_ids) cudaFree(ctx->d_ids);
    ctx->d_pos = np; ctx->d_wi = nw; ctx->d_alpha = na; ctx->d_ids = ni; ctx->cap_photons = cap;
    return PV_OK;
}
}  // extern "C"
int pvi_reserve_photons(pv_ctx *ctx, uint64_t n) { return reserve_photons(ctx, n); }
extern "C" {

__global__ void iota_ids_kernel(uint64_t *ids, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = i;
}
// ABI planes carry 30 floats per photon; the context keeps 32 (one 128-byte line)
__global__ void alpha_30_to_32_kernel(const float *__restrict__ a30, float *__restrict__ a32, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 32) return;
    uint64_t ph = i >> 5; uint32_t b = (uint32_t)(i & 31);
    a32[i] = b < PV_NSPEC ? a30[ph * PV_NSPEC + b] : 0.f;
}
__global__ void alpha_32_to_30_kernel(const float *__restrict__ a32, float *__restrict__ a30, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * PV_NSPEC) return;
    uint64_t ph = i / PV_NSPEC; uint32_t b = (uint32_t)(i % PV_NSPEC);
    a30[i] = a32[ph * 32 + b];
}

static int set_photons_impl(pv_ctx *ctx, const float *pos, const float *wi, const float *alpha, uint64_t n, cudaMemcpyKind kind) {
    if (n && (!pos || !wi || !alpha)) { ctx->err = "pv_set_photons: null plane"; return PV_EINVAL; }
    ctx->built = false;
    ctx->n_photons = 0;
    int rc = reserve_photons(ctx, n); if (rc) return rc;
    if (n) {
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_pos, pos, n * 3 * sizeof(float), kind, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->d_wi, wi, n * 3 * sizeof(float), kind, ctx->stream));
        const float *a30 = alpha;
        if (kind == cudaMemcpyHostToDevice) {
            int rc2 = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, n * PV_NSPEC * sizeof(float)); if (rc2) return rc2;
            PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->io, alpha, n * PV_NSPEC * sizeof(float), kind, ctx->stream));
            a30 = (const float *)ctx->io;
        }
        alpha_30_to_32_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(a30, ctx->d_alpha, n);
        iota_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_ids, n);
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->n_photons = n;
    return PV_OK;
}
int pv_set_photons(pv_ctx *ctx, const float *pos, const float *wi, const float *alpha, uint64_t n) {
    LOCK(ctx);
    return set_photons_impl(ctx, pos, wi, alpha, n, cudaMemcpyHostToDevice);
}
int pv_set_photons_dev(pv_ctx *ctx, const float *pos, const float *wi, const float *alpha, uint64_t n) {
    LOCK(ctx);
    return set_photons_impl(ctx, pos, wi, alpha, n, cudaMemcpyDeviceToDevice);
}
static int get_photons_impl(pv_ctx *ctx, float *pos, float *wi, float *alpha, uint64_t *ids, uint64_t capacity, uint64_t *n, cudaMemcpyKind kind) {
    uint64_t m = std::min<uint64_t>(capacity, ctx->n_photons);
    if (n) *n = m;
    if (!m) return PV_OK;
    if (pos) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(pos, ctx->d_pos, m * 3 * sizeof(float), kind, ctx->stream));
    if (wi) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(wi, ctx->d_wi, m * 3 * sizeof(float), kind, ctx->stream));
    if (alpha) {
        if (kind == cudaMemcpyDeviceToHost) {
            int rc2 = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, m * PV_NSPEC * sizeof(float)); if (rc2) return rc2;
            alpha_32_to_30_kernel<<<(unsigned)((m * PV_NSPEC + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_alpha, (float *)ctx->io, m);
            PV_CUDA_CHECK(ctx, cudaMemcpyAsync(alpha, ctx->io, m * PV_NSPEC * sizeof(float), kind, ctx->stream));
        } else {
            alpha_32_to_30_kernel<<<(unsigned)((m * PV_NSPEC + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_alpha, alpha, m);
        }
    }
    if (ids) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ids, ctx->d_ids, m * sizeof(uint64_t), kind, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_get_photons(pv_ctx *ctx, float *pos, float *wi, float *alpha, uint64_t *ids, uint64_t capacity, uint64_t *n) {
    LOCK(ctx);
    return get_photons_impl(ctx, pos, wi, alpha, ids, capacity, n, cudaMemcpyDeviceToHost);
}
int pv_get_photons_dev(pv_ctx *ctx, float *pos, float *wi, float *alpha, uint64_t *ids, uint64_t capacity, uint64_t *n) {
    LOCK(ctx);
    return get_photons_impl(ctx, pos, wi, alpha, ids, capacity, n, cudaMemcpyDeviceToDevice);
}
int pv_photon_count(pv_ctx *ctx, uint64_t *n) {
    LOCK(ctx);
    if (!n) { ctx->err = "pv_photon_count: null out"; return PV_EINVAL; }
    *n = ctx->n_photons;
    return PV_OK;
}
int pv_build(pv_ctx *ctx, float maxdist, uint32_t nused) {
    LOCK(ctx);
    return pvi_build(ctx, maxdist, nused);
}

// host-pointer staging helpers
static int stage_in(pv_ctx *ctx, void **buf, size_t *cap, const void *src, size_t bytes) {
    int rc = pv_ensure(ctx, buf, cap, bytes); if (rc) return rc;
    if (bytes) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(*buf, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PV_OK;
}

int pv_knn(pv_ctx *ctx, const float *pts, uint64_t n, uint32_t k, float r2, uint32_t *idx, float *d2, uint32_t *nfound) {
    LOCK(ctx);
    if (n && (!pts || !idx || !d2 || !nfound)) { ctx->err = "pv_knn: null pointer"; return PV_EINVAL; }
    if (!ctx->built) { ctx->err = "pv_knn: photon map not built (call pv_build)"; return PV_ESTATE; }
    if (!n || !k) return PV_OK;
    int rc = stage_in(ctx, &ctx->io, &ctx->io_bytes, pts, n * 3 * sizeof(float)); if (rc) return rc;
    size_t ob = n * k * 8 + n * 4;
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, ob); if (rc) return rc;
    uint32_t *d_idx = (uint32_t *)ctx->io2; float *d_d2 = (float *)(d_idx + n * k); uint32_t *d_nf = (uint32_t *)(d_d2 + n * k);
    rc = pvi_knn(ctx, (const float *)ctx->io, n, k, r2, d_idx, d_d2, d_nf); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(idx, d_idx, n * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d2, d_d2, n * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(nfound, d_nf, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}

int pv_lphoton(pv_ctx *ctx, const float *pts, const float *w, uint64_t n, uint32_t nused, float maxdist, float *L) {
    LOCK(ctx);
    if (n && (!pts || !w || !L)) { ctx->err = "pv_lphoton: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, n * 6 * sizeof(float)); if (rc) return rc;
    float *d_pts = (float *)ctx->io, *d_w = d_pts + n * 3;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pts, pts, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_w, w, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, n * PV_NSPEC * sizeof(float)); if (rc) return rc;
    rc = pvi_lphoton(ctx, d_pts, d_w, n, nused, maxdist, (float *)ctx->io2); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(L, ctx->io2, n * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}

int pv_intersect(pv_ctx *ctx, const pv_ray *rays, uint64_t n, uint32_t *prim, float *t) {
    LOCK(ctx);
    if (n && (!rays || !prim || !t)) { ctx->err = "pv_intersect: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    int rc = stage_in(ctx, &ctx->io, &ctx->io_bytes, rays, n * sizeof(pv_ray)); if (rc) return rc;
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, n * 8); if (rc) return rc;
    uint32_t *d_prim = (uint32_t *)ctx->io2; float *d_t = (float *)(d_prim + n);
    rc = pvi_intersect(ctx, (const pv_ray *)ctx->io, n, d_prim, d_t); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(prim, d_prim, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(t, d_t, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_build_bvh(pv_ctx *ctx, const float *prim_bounds, uint32_t n_prims, uint32_t max_prims_in_node, pv_bvh_node *nodes, uint32_t nodes_cap,
                 uint32_t *n_nodes, uint32_t *prim_order, float *device_ms) {
    LOCK(ctx);
    if (!n_nodes || (n_prims && (!prim_bounds || !nodes || !prim_order))) { ctx->err = "pv_build_bvh: null pointer"; return PV_EINVAL; }
    if (max_prims_in_node < 1 || max_prims_in_node > 255) {       // BVHAccel clamps to 255 too (bvh.cpp:198): nPrimitives is a byte
        ctx->err = "pv_build_bvh: max_prims_in_node must be in [1, 255]"; return PV_EINVAL;
    }
    return pvi_build_bvh(ctx, prim_bounds, n_prims, max_prims_in_node, nodes, nodes_cap, n_nodes, prim_order, device_ms);
}
int pv_occluded(pv_ctx *ctx, const pv_ray *rays, uint64_t n, uint8_t *hit) {
    LOCK(ctx);
    if (n && (!rays || !hit)) { ctx->err = "pv_occluded: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    int rc = stage_in(ctx, &ctx->io, &ctx->io_bytes, rays, n * sizeof(pv_ray)); if (rc) return rc;
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, n); if (rc) return rc;
    rc = pvi_occluded(ctx, (const pv_ray *)ctx->io, n, (uint8_t *)ctx->io2); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(hit, ctx->io2, n, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_transmittance(pv_ctx *ctx, const pv_ray *rays, uint64_t n, float step, const float *offset_u, float *T) {
    LOCK(ctx);
    if (n && (!rays || !T)) { ctx->err = "pv_transmittance: null pointer"; return PV_EINVAL; }
    if (!(step > 0.f)) { ctx->err = "pv_transmittance: step must be > 0"; return PV_EINVAL; }
    if (!n) return PV_OK;
    size_t rb = n * sizeof(pv_ray), ub = n * sizeof(float);
    int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, rb + ub); if (rc) return rc;
    pv_ray *d_rays = (pv_ray *)ctx->io; float *d_u = (float *)((char *)ctx->io + rb);
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_rays, rays, rb, cudaMemcpyHostToDevice, ctx->stream));
    if (offset_u) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_u, offset_u, ub, cudaMemcpyHostToDevice, ctx->stream));
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, n * PV_NSPEC * sizeof(float)); if (rc) return rc;
    rc = pvi_transmittance(ctx, d_rays, n, step, offset_u ? d_u : nullptr, (float *)ctx->io2); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(T, ctx->io2, n * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}

int pv_gather_dev(pv_ctx *ctx, const pv_ray *rays, uint64_t n, const pv_gather_params *params, float *L, float *T) {
    LOCK(ctx);
    if (!params || (n && (!rays || !L || !T))) { ctx->err = "pv_gather_dev: null pointer"; return PV_EINVAL; }
    int rc = pvi_gather(ctx, rays, n, params, L, T); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
// Host-pointer Li: the frame is cut into slices; the host->device copy of slice s+1 and the device->host copy of
// slice s-1 run on their own streams while slice s is gathered, so only the first upload and the last download are
// exposed (the result is 240 B per ray, ~0.5 GB for a 1080p frame).  Asynchronous only if the caller's buffers are
// pinned; with pageable memory CUDA serialises the copies and the result is the same.
int pv_gather(pv_ctx *ctx, const pv_ray *rays, uint64_t n, const pv_gather_params *params, float *L, float *T) {
    LOCK(ctx);
    if (!params || (n && (!rays || !L || !T))) { ctx->err = "pv_gather: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, n * sizeof(pv_ray)); if (rc) return rc;
    const size_t sb = n * PV_NSPEC * sizeof(float);
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, 2 * sb); if (rc) return rc;
    pv_ray *d_rays = (pv_ray *)ctx->io;
    float *d_L = (float *)ctx->io2, *d_T = d_L + n * PV_NSPEC;
    if (!ctx->copy_in) {
        PV_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        PV_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    }
    const uint64_t min_slice = 1ull << 18;
    int nslices = (int)std::max<uint64_t>(1, std::min<uint64_t>(4, n / min_slice));
    if (const char *e = getenv("PV_GATHER_SLICES")) nslices = std::max(1, std::min(PV_GATHER_MAX_SLICES, atoi(e)));     // tuning knob
    // Every path out of this function (the CUDA checks below return early on an error) first waits for the copies that read or
    // write the CALLER's buffers, then releases the per-slice events.
    struct SliceEvents {
        cudaStream_t in, out;
        cudaEvent_t in_done[PV_GATHER_MAX_SLICES] = {}, comp_done[PV_GATHER_MAX_SLICES] = {};
        ~SliceEvents() {
            cudaStreamSynchronize(in); cudaStreamSynchronize(out);
            for (int s = 0; s < PV_GATHER_MAX_SLICES; ++s) { if (in_done[s]) cudaEventDestroy(in_done[s]); if (comp_done[s]) cudaEventDestroy(comp_done[s]); }
        }
    } evs{ctx->copy_in, ctx->copy_out};
    cudaEvent_t *in_done = evs.in_done, *comp_done = evs.comp_done;
    for (int s = 0; s < nslices; ++s) {
        PV_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&in_done[s], cudaEventDisableTiming));
        PV_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&comp_done[s], cudaEventDisableTiming));
    }
    auto lo = [&](int s) { return n * (uint64_t)s / (uint64_t)nslices; };
    // the device buffers of an earlier call may still be read by work on ctx->stream
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < nslices; ++s) {
        const uint64_t a = lo(s), b = lo(s + 1);
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_rays + a, rays + a, (b - a) * sizeof(pv_ray), cudaMemcpyHostToDevice, ctx->copy_in));
        PV_CUDA_CHECK(ctx, cudaEventRecord(in_done[s], ctx->copy_in));
    }
    float ms = 0.f, march_ms = 0.f, phase_ms[4] = {0.f, 0.f, 0.f, 0.f};
    rc = PV_OK;
    for (int s = 0; s < nslices && rc == PV_OK; ++s) {
        const uint64_t a = lo(s), b = lo(s + 1);
        pv_gather_params p = *params; p.ray_index_base = params->ray_index_base + a;
        PV_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, in_done[s], 0));
        rc = pvi_gather(ctx, d_rays + a, b - a, &p, d_L + a * PV_NSPEC, d_T + a * PV_NSPEC);
        if (rc) break;
        ms += ctx->last_ms; march_ms += ctx->last_march_ms;
        for (int i = 0; i < 4; ++i) phase_ms[i] += ctx->phase_ms[i];
        PV_CUDA_CHECK(ctx, cudaEventRecord(comp_done[s], ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_out, comp_done[s], 0));
        const size_t ob = (b - a) * PV_NSPEC * sizeof(float);
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(L + a * PV_NSPEC, d_L + a * PV_NSPEC, ob, cudaMemcpyDeviceToHost, ctx->copy_out));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(T + a * PV_NSPEC, d_T + a * PV_NSPEC, ob, cudaMemcpyDeviceToHost, ctx->copy_out));
    }
    cudaStreamSynchronize(ctx->copy_in);
    cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->copy_out);
    if (rc) return rc;
    if (e != cudaSuccess) { ctx->err = std::string("pv_gather: ") + cudaGetErrorString(e); return PV_ECUDA; }
    ctx->last_ms = ms; ctx->last_march_ms = march_ms;
    for (int i = 0; i < 4; ++i) ctx->phase_ms[i] = phase_ms[i];
    return PV_OK;
}
int pv_volume_li_dev(pv_ctx *ctx, int integrator, const pv_ray *rays, uint64_t n, const pv_gather_params *params, float *L, float *T) {
    LOCK(ctx);
    if (!params || (n && (!rays || !L || !T))) { ctx->err = "pv_volume_li_dev: null pointer"; return PV_EINVAL; }
    int rc = pvi_volume_li(ctx, integrator, rays, n, params, L, T); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_volume_li(pv_ctx *ctx, int integrator, const pv_ray *rays, uint64_t n, const pv_gather_params *params, float *L, float *T) {
    LOCK(ctx);
    if (!params || (n && (!rays || !L || !T))) { ctx->err = "pv_volume_li: null pointer"; return PV_EINVAL; }
    const size_t rb = n * sizeof(pv_ray), sb = n * PV_NSPEC * sizeof(float);
    int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, rb); if (rc) return rc;
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, 2 * sb); if (rc) return rc;
    pv_ray *d_rays = (pv_ray *)ctx->io;
    float *d_L = (float *)ctx->io2, *d_T = d_L + n * PV_NSPEC;
    if (n) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_rays, rays, rb, cudaMemcpyHostToDevice, ctx->stream));
    rc = pvi_volume_li(ctx, integrator, d_rays, n, params, d_L, d_T); if (rc) return rc;
    if (n) {
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(L, d_L, sb, cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(T, d_T, sb, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
// Li of n rays with a stream index of its own per ray: what a caller that collects rays from many threads into one call needs for its
// results not to depend on who shared a call with whom.  integrator < 0: PhotonVolumeIntegrator::Li, else pv_volume_li's.
static int gather_indexed(pv_ctx *ctx, int integrator, const pv_ray *rays, const uint64_t *ray_index, uint64_t n, const pv_gather_params *params,
                          float *L, float *T) {
    if (!params || (n && (!rays || !ray_index || !L || !T))) { ctx->err = "pv_gather_indexed: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    const size_t rb = n * sizeof(pv_ray), sb = n * PV_NSPEC * sizeof(float);
    int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, rb); if (rc) return rc;
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, 2 * sb); if (rc) return rc;
    rc = pv_ensure(ctx, &ctx->io3, &ctx->io3_bytes, n * sizeof(uint64_t)); if (rc) return rc;
    pv_ray *d_rays = (pv_ray *)ctx->io;
    float *d_L = (float *)ctx->io2, *d_T = d_L + n * PV_NSPEC;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_rays, rays, rb, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->io3, ray_index, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->d_ray_index = (const uint64_t *)ctx->io3; ctx->ray_index_rays = d_rays;
    rc = integrator < 0 ? pvi_gather(ctx, d_rays, n, params, d_L, d_T) : pvi_volume_li(ctx, integrator, d_rays, n, params, d_L, d_T);
    ctx->d_ray_index = nullptr; ctx->ray_index_rays = nullptr;
    if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(L, d_L, sb, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(T, d_T, sb, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_gather_indexed(pv_ctx *ctx, const pv_ray *rays, const uint64_t *ray_index, uint64_t n, const pv_gather_params *params, float *L, float *T) {
    LOCK(ctx);
    return gather_indexed(ctx, -1, rays, ray_index, n, params, L, T);
}
int pv_volume_li_indexed(pv_ctx *ctx, int integrator, const pv_ray *rays, const uint64_t *ray_index, uint64_t n, const pv_gather_params *params,
                         float *L, float *T) {
    LOCK(ctx);
    if (integrator < 0) { ctx->err = "pv_volume_li_indexed: bad integrator"; return PV_EINVAL; }
    return gather_indexed(ctx, integrator, rays, ray_index, n, params, L, T);
}
int pv_gather_stats_get(pv_ctx *ctx, pv_gather_stats *out, int reset) {
    LOCK(ctx);
    if (!out) { ctx->err = "pv_gather_stats_get: null out"; return PV_EINVAL; }
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpy(out, ctx->d_stats, sizeof(*out), cudaMemcpyDeviceToHost));
    if (reset) PV_CUDA_CHECK(ctx, cudaMemset(ctx->d_stats, 0, sizeof(*out)));
    return PV_OK;
}
int pv_last_kernel_ms(pv_ctx *ctx, float *ms) {
    LOCK(ctx);
    if (!ms) { ctx->err = "pv_last_kernel_ms: null out"; return PV_EINVAL; }
    *ms = ctx->last_ms;
    return PV_OK;
}
int pv_comm_unique_id(uint8_t *id) {
    if (!id) { g_err = "pv_comm_unique_id: null id"; return PV_EINVAL; }
    return pvi_comm_unique_id(id, &g_err);
}
int pv_comm_init(pv_ctx *ctx, const uint8_t *id, int rank, int world) {
    LOCK(ctx);
    if (!id) { ctx->err = "pv_comm_init: null id"; return PV_EINVAL; }
    return pvi_comm_init(ctx, id, rank, world);
}
int pv_comm_init_all(pv_ctx **ctxs, int n) {
    if (!ctxs || n < 1) { g_err = "pv_comm_init_all: no contexts"; return PV_EINVAL; }
    for (int i = 0; i < n; ++i) if (!ctxs[i]) { g_err = "pv_comm_init_all: null context"; return PV_EINVAL; }
    int rc = pvi_comm_init_all(ctxs, n);
    cudaSetDevice(ctxs[0]->device);
    return rc;
}
int pv_comm_destroy(pv_ctx *ctx) {
    LOCK(ctx);
    return pvi_comm_destroy(ctx);
}
int pv_allgather_photons(pv_ctx *ctx, int renumber, float *collective_ms) {
    LOCK(ctx);
    return pvi_allgather_photons(ctx, renumber, collective_ms);
}
int pv_broadcast_photons(pv_ctx **ctxs, int n, int src, float *collective_ms) {
    if (!ctxs || n < 1 || src < 0 || src >= n) { g_err = "pv_broadcast_photons: bad arguments"; return PV_EINVAL; }
    for (int i = 0; i < n; ++i) if (!ctxs[i]) { g_err = "pv_broadcast_photons: null context"; return PV_EINVAL; }
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int i = 0; i < n; ++i) locks.emplace_back(ctxs[i]->mu);
    int rc = pvi_broadcast_photons(ctxs, n, src, collective_ms);
    cudaSetDevice(ctxs[src]->device);
    return rc;
}
int pv_launch_count(pv_ctx *ctx, uint64_t *n) {
    LOCK(ctx);
    if (!n) { ctx->err = "pv_launch_count: null out"; return PV_EINVAL; }
    *n = ctx->launches;
    return PV_OK;
}
int pv_last_phase_ms(pv_ctx *ctx, float ms[4]) {
    LOCK(ctx);
    if (!ms) { ctx->err = "pv_last_phase_ms: null out"; return PV_EINVAL; }
    for (int i = 0; i < 4; ++i) ms[i] = ctx->phase_ms[i];
    return PV_OK;
}
int pv_last_march_ms(pv_ctx *ctx, float *ms) {
    LOCK(ctx);
    if (!ms) { ctx->err = "pv_last_march_ms: null out"; return PV_EINVAL; }
    *ms = ctx->last_march_ms;
    return PV_OK;
}
int pv_shoot(pv_ctx *ctx, uint64_t n_volume_wanted, const pv_shoot_params *params, pv_shoot_stats *stats) {
    LOCK(ctx);
    if (!params) { ctx->err = "pv_shoot: null params"; return PV_EINVAL; }
    return pvi_shoot(ctx, n_volume_wanted, params, stats);
}
int pv_shoot_blocks(pv_ctx *ctx, uint64_t first_block, uint32_t n_blocks, const pv_shoot_params *params, uint32_t *counts,
                    pv_shoot_stats *stats) {
    LOCK(ctx);
    if (!params || (n_blocks && !counts)) { ctx->err = "pv_shoot_blocks: null pointer"; return PV_EINVAL; }
    return pvi_shoot_blocks(ctx, first_block, n_blocks, params, counts, stats);
}
int pv_shoot_finish(pv_ctx *ctx, uint64_t last_block) {
    LOCK(ctx);
    return pvi_shoot_finish(ctx, last_block);
}

// ---- surface photon maps
int pv_shoot_maps(pv_ctx *ctx, const pv_maps_params *maps, const pv_shoot_params *params, pv_maps_stats *stats) {
    LOCK(ctx);
    if (!maps || !params) { ctx->err = "pv_shoot_maps: null params"; return PV_EINVAL; }
    return pvi_shoot_maps(ctx, maps, params, nullptr, nullptr, stats);
}
int pv_shoot_maps_ranks(pv_ctx *ctx, const pv_maps_params *maps, const pv_shoot_params *params, pv_allreduce_u32_fn allreduce, void *user,
                        pv_maps_stats *stats) {
    LOCK(ctx);
    if (!maps || !params) { ctx->err = "pv_shoot_maps_ranks: null params"; return PV_EINVAL; }
    return pvi_shoot_maps(ctx, maps, params, allreduce, user, stats);
}
int pv_get_map_photons(pv_ctx *ctx, int map, float *pos, float *wi, float *alpha, uint64_t *ids, uint64_t capacity, uint64_t *n) {
    LOCK(ctx);
    if (map == PV_MAP_VOLUME) return get_photons_impl(ctx, pos, wi, alpha, ids, capacity, n, cudaMemcpyDeviceToHost);
    if (map < PV_MAP_CAUSTIC || map > PV_MAP_RADIANCE) { ctx->err = "pv_get_map_photons: bad map"; return PV_EINVAL; }
    const PhotonSet &s = ctx->surf[map - 1];
    uint64_t m = std::min<uint64_t>(capacity, s.n);
    if (n) *n = m;
    if (!m) return PV_OK;
    if (pos) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(pos, s.pos, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (wi) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(wi, s.wi, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (alpha) {
        int rc2 = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, m * PV_NSPEC * sizeof(float)); if (rc2) return rc2;
        alpha_32_to_30_kernel<<<(unsigned)((m * PV_NSPEC + 255) / 256), 256, 0, ctx->stream>>>(s.alpha, (float *)ctx->io, m);
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(alpha, ctx->io, m * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (ids) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ids, s.ids, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_set_map_photons(pv_ctx *ctx, int map, const float *pos, const float *wi, const float *alpha, uint64_t n) {
    LOCK(ctx);
    if (map < PV_MAP_CAUSTIC || map > PV_MAP_RADIANCE) { ctx->err = "pv_set_map_photons: map must be PV_MAP_CAUSTIC..PV_MAP_RADIANCE"; return PV_EINVAL; }
    if (n && (!pos || !wi || !alpha)) { ctx->err = "pv_set_map_photons: null plane"; return PV_EINVAL; }
    PhotonSet &s = ctx->surf[map - 1];
    s.n = 0; ctx->rad_valid = false;
    if (ctx->map_which == map) ctx->built = false;
    int rc = pvi_reserve_set(ctx, &s, n); if (rc) return rc;
    if (n) {
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(s.pos, pos, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(s.wi, wi, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, n * PV_NSPEC * sizeof(float)); if (rc) return rc;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->io, alpha, n * PV_NSPEC * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        alpha_30_to_32_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>((const float *)ctx->io, s.alpha, n);
        iota_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(s.ids, n);
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    s.n = n;
    return PV_OK;
}
int pv_set_radiance_lo(pv_ctx *ctx, const float *Lo, uint64_t n) {
    LOCK(ctx);
    ctx->rad_valid = false;
    if (n != ctx->surf[3].n) { ctx->err = "pv_set_radiance_lo: n differs from the number of radiance photons"; return PV_EINVAL; }
    if (n && !Lo) { ctx->err = "pv_set_radiance_lo: null pointer"; return PV_EINVAL; }
    if (n) {
        if (ctx->rad_Lo_cap < n) {
            if (ctx->rad_Lo) cudaFree(ctx->rad_Lo);
            ctx->rad_Lo = nullptr; ctx->rad_Lo_cap = 0;
            PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ctx->rad_Lo, n * 32 * sizeof(float)));
            ctx->rad_Lo_cap = n;
        }
        int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, n * PV_NSPEC * sizeof(float)); if (rc) return rc;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->io, Lo, n * PV_NSPEC * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        alpha_30_to_32_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>((const float *)ctx->io, ctx->rad_Lo, n);
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->rad_valid = true;
    return PV_OK;
}
int pv_select_map(pv_ctx *ctx, int map, float maxdist, uint32_t nused) {
    LOCK(ctx);
    return pvi_build_map(ctx, map, maxdist, nused);
}
int pv_surface_lphoton(pv_ctx *ctx, const float *pts, const float *nf, uint64_t n, uint32_t n_lookup, float max_dist2, uint64_t n_paths,
                       float *Lr, float *Lt) {
    LOCK(ctx);
    if (n && (!pts || !nf || !Lr || !Lt)) { ctx->err = "pv_surface_lphoton: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, n * 6 * sizeof(float)); if (rc) return rc;
    float *d_pts = (float *)ctx->io, *d_nf = d_pts + n * 3;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pts, pts, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_nf, nf, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, 2 * n * PV_NSPEC * sizeof(float)); if (rc) return rc;
    float *d_Lr = (float *)ctx->io2, *d_Lt = d_Lr + n * PV_NSPEC;
    rc = pvi_surface_lphoton(ctx, d_pts, d_nf, n, n_lookup, max_dist2, n_paths, d_Lr, d_Lt); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(Lr, d_Lr, n * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(Lt, d_Lt, n * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_radiance_nearest(pv_ctx *ctx, const float *pts, const float *normals, uint64_t n, uint32_t *idx, float *Lo) {
    LOCK(ctx);
    if (n && (!pts || !normals || !idx)) { ctx->err = "pv_radiance_nearest: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    int rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, n * 6 * sizeof(float)); if (rc) return rc;
    float *d_pts = (float *)ctx->io, *d_n = d_pts + n * 3;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pts, pts, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_n, normals, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, n * sizeof(uint32_t) + n * PV_NSPEC * sizeof(float)); if (rc) return rc;
    float *d_Lo = (float *)ctx->io2; uint32_t *d_idx = (uint32_t *)(d_Lo + n * PV_NSPEC);
    rc = pvi_radiance_nearest(ctx, d_pts, d_n, n, d_idx, Lo ? d_Lo : nullptr); if (rc) return rc;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(idx, d_idx, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (Lo) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(Lo, d_Lo, n * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pv_final_gather(pv_ctx *ctx, const pv_ray *rays, uint64_t n, float step, uint64_t seed, uint64_t index_base, float *Lindir, uint32_t *idx) {
    LOCK(ctx);
    if (n && (!rays || !Lindir)) { ctx->err = "pv_final_gather: null pointer"; return PV_EINVAL; }
    if (!n) return PV_OK;
    // slices bound the device staging (30 floats out per ray)
    const uint64_t slice = 4ull << 20;
    for (uint64_t a = 0; a < n; a += slice) {
        const uint64_t m = std::min<uint64_t>(slice, n - a);
        int rc = stage_in(ctx, &ctx->io, &ctx->io_bytes, rays + a, m * sizeof(pv_ray)); if (rc) return rc;
        rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, m * PV_NSPEC * sizeof(float) + m * sizeof(uint32_t)); if (rc) return rc;
        float *d_L = (float *)ctx->io2; uint32_t *d_idx = (uint32_t *)(d_L + m * PV_NSPEC);
        rc = pvi_final_gather(ctx, (const pv_ray *)ctx->io, m, step, seed, index_base + a, d_L, d_idx); if (rc) return rc;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(Lindir + a * PV_NSPEC, d_L, m * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (idx) PV_CUDA_CHECK(ctx, cudaMemcpyAsync(idx + a, d_idx, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PV_OK;
}
int pv_radiance_photons(pv_ctx *ctx, uint32_t n_lookup, float max_dist2, const uint64_t *path_counts, float *Lo, uint64_t capacity, uint64_t *n) {
    LOCK(ctx);
    const uint64_t own[3] = {ctx->map_paths[2], ctx->map_paths[1], ctx->map_paths[0]};      // direct, indirect, caustic
    int rc = pvi_radiance(ctx, n_lookup, max_dist2, path_counts ? path_counts : own); if (rc) return rc;
    uint64_t m = std::min<uint64_t>(capacity, ctx->surf[3].n);
    if (n) *n = m;
    if (!m || !Lo) return PV_OK;
    rc = pv_ensure(ctx, &ctx->io, &ctx->io_bytes, m * PV_NSPEC * sizeof(float)); if (rc) return rc;
    alpha_32_to_30_kernel<<<(unsigned)((m * PV_NSPEC + 255) / 256), 256, 0, ctx->stream>>>(ctx->rad_Lo, (float *)ctx->io, m);
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(Lo, ctx->io, m * PV_NSPEC * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}

}  // extern "C"
