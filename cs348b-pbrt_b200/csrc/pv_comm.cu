// pv_comm.cu -- NCCL behind the C ABI: the photon map's replication step of SURVEY.md 8(e).
//
// The reference has no multi-process path (its parallelism is pthread tasks in one address space, core/parallel.cpp:728-737,
// and the map is one shared vector, core/photonshooter.cpp:333-337).  Here emission is sharded by 4096-path block over the
// GPUs of a box and every GPU gathers from its own copy of the map, so the map is replicated ONCE per frame:
//
//   pv_allgather_photons   one process per GPU (pv_comm_init): every rank contributes its photons, every rank ends with the
//                          union.  One ncclAllGather per SoA plane (pos, wi, alpha32, ids), grouped, straight from the
//                          shooter's output planes into the planes pv_build reads -- no host staging, no padded staging
//                          tensors, no index_select: ranks' segments sit at stride max-count in the receive planes, the
//                          few slots of slack between them carry id = ~0 and drop out in the one pass that follows anyway,
//                          the radix sort by photon id that makes the set independent of the number of ranks.
//   pv_broadcast_photons   one process driving several GPUs (pv_comm_init_all; the drop-in's PV_DEVICES): the set of one
//                          context is broadcast to the others.
//
// libnccl is opened at run time (dlopen "libnccl.so.2"): libpv.so itself has no NCCL dependency, a single-GPU host never
// loads it, and inside a process that already carries an NCCL (PyTorch's) the same library instance is used.
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <vector>
#include "pv_ctx.h"

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    std::string err;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

const NcclApi *nccl_api(std::string *err) {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.lib) return &g_nccl;
    if (!g_nccl.err.empty()) { *err = g_nccl.err; return nullptr; }
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl.err = std::string("NCCL not available: ") + dlerror(); *err = g_nccl.err; return nullptr; }
    bool ok = true;
    auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) { ok = false; g_nccl.err = std::string("NCCL symbol missing: ") + name; } return p; };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.CommGetAsyncError = (decltype(g_nccl.CommGetAsyncError))sym("ncclCommGetAsyncError");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))sym("ncclBroadcast");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    if (!ok) { dlclose(h); *err = g_nccl.err; return nullptr; }
    g_nccl.lib = h;
    return &g_nccl;
}
}  // namespace

#define PV_NCCL_CHECK(ctx, api, call)                                                                         \
    do {                                                                                                      \
        ncclResult_t r__ = (call);                                                                            \
        if (r__ != ncclSuccess) { (ctx)->err = std::string(#call) + ": " + (api)->GetErrorString(r__); return PV_ECUDA; } \
    } while (0)

// the communicator's asynchronous error state into pv_last_error
static int comm_async_check(pv_ctx *ctx, const NcclApi *api, const char *where) {
    ncclResult_t async = ncclSuccess;
    ncclResult_t r = api->CommGetAsyncError((ncclComm_t)ctx->comm, &async);
    if (r != ncclSuccess || (async != ncclSuccess && async != ncclInProgress)) {
        ctx->err = std::string(where) + ": NCCL " + api->GetErrorString(r != ncclSuccess ? r : async);
        return PV_ECUDA;
    }
    return PV_OK;
}

__global__ void renumber_ids_kernel(uint64_t *ids, uint64_t n, uint64_t base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = base + i;
}

int pvi_comm_unique_id(uint8_t *id, std::string *err) {
    const NcclApi *api = nccl_api(err);
    if (!api) return PV_ESTATE;
    ncclUniqueId u;
    ncclResult_t r = api->GetUniqueId(&u);
    if (r != ncclSuccess) { *err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return PV_ECUDA; }
    static_assert(sizeof(u) == PV_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, sizeof(u));
    return PV_OK;
}

int pvi_comm_init(pv_ctx *ctx, const uint8_t *id, int rank, int world) {
    if (ctx->comm) { ctx->err = "pv_comm_init: the context already has a communicator (pv_comm_destroy first)"; return PV_ESTATE; }
    if (world < 1 || rank < 0 || rank >= world) { ctx->err = "pv_comm_init: bad rank / world"; return PV_EINVAL; }
    const NcclApi *api = nccl_api(&ctx->err);
    if (!api) return PV_ESTATE;
    ncclUniqueId u; memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    PV_NCCL_CHECK(ctx, api, api->CommInitRank(&c, world, u, rank));
    ctx->comm = c; ctx->comm_rank = rank; ctx->comm_world = world;
    return PV_OK;
}

int pvi_comm_init_all(pv_ctx **ctxs, int n) {
    pv_ctx *c0 = ctxs[0];
    const NcclApi *api = nccl_api(&c0->err);
    if (!api) return PV_ESTATE;
    std::vector<int> devs(n); std::vector<ncclComm_t> comms(n, nullptr);
    for (int i = 0; i < n; ++i) {
        if (ctxs[i]->comm) { c0->err = "pv_comm_init_all: a context already has a communicator"; return PV_ESTATE; }
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; ++j) if (devs[j] == devs[i]) { c0->err = "pv_comm_init_all: two contexts on one device"; return PV_EINVAL; }
    }
    PV_NCCL_CHECK(c0, api, api->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) { ctxs[i]->comm = comms[i]; ctxs[i]->comm_rank = i; ctxs[i]->comm_world = n; }
    return PV_OK;
}

int pvi_comm_destroy(pv_ctx *ctx) {
    if (!ctx->comm) return PV_OK;
    std::string e;
    const NcclApi *api = nccl_api(&e);
    if (api) { cudaStreamSynchronize(ctx->stream); api->CommDestroy((ncclComm_t)ctx->comm); }
    ctx->comm = nullptr; ctx->comm_world = 0; ctx->comm_rank = 0;
    return PV_OK;
}

int pvi_sort_photons_by_id(pv_ctx *ctx);     // pv_shoot.cu

int pvi_allgather_photons(pv_ctx *ctx, int renumber, float *collective_ms) {
    if (collective_ms) *collective_ms = 0.f;
    if (!ctx->comm) { ctx->err = "pv_allgather_photons: no communicator (pv_comm_init first)"; return PV_ESTATE; }
    const NcclApi *api = nccl_api(&ctx->err);
    if (!api) return PV_ESTATE;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const int W = ctx->comm_world, rank = ctx->comm_rank;
    ctx->built = false;
    // 1. how many photons each rank holds
    int rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, (size_t)(W + 1) * sizeof(uint64_t)); if (rc) return rc;
    uint64_t *d_cnt = (uint64_t *)ctx->io2;
    const uint64_t n_local = ctx->n_photons;
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cnt + W, &n_local, sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    PV_NCCL_CHECK(ctx, api, api->AllGather(d_cnt + W, d_cnt, 1, ncclUint64, comm, ctx->stream));
    std::vector<uint64_t> counts(W);
    PV_CUDA_CHECK(ctx, cudaMemcpyAsync(counts.data(), d_cnt, (size_t)W * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    rc = comm_async_check(ctx, api, "pv_allgather_photons"); if (rc) return rc;
    uint64_t mx = 0, total = 0, before = 0;
    for (int r = 0; r < W; ++r) { mx = std::max(mx, counts[r]); if (r < rank) before += counts[r]; total += counts[r]; }
    if (total == 0) return PV_OK;
    if (total > 0xFFFFFFF0ull || (uint64_t)W * mx > 0xFFFFFFF0ull) { ctx->err = "pv_allgather_photons: too many photons for 32-bit indices"; return PV_EINVAL; }
    // 2. send planes hold mx records: the slack behind this rank's photons is marked id = ~0
    rc = pvi_reserve_photons(ctx, mx); if (rc) return rc;
    if (renumber && n_local) {                            // injected sets: photon i of the union keeps the index a single-rank set gives it
        renumber_ids_kernel<<<(unsigned)((n_local + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_ids, n_local, before);
        PV_CUDA_CHECK(ctx, cudaGetLastError());
    }
    if (mx > n_local) PV_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_ids + n_local, 0xFF, (mx - n_local) * sizeof(uint64_t), ctx->stream));
    // 3. receive planes: rank r's segment at r * mx
    const uint64_t cap = (uint64_t)W * mx;
    float *np = nullptr, *nw = nullptr, *na = nullptr; uint64_t *ni = nullptr;
    auto fail = [&](int code) { if (np) cudaFree(np); if (nw) cudaFree(nw); if (na) cudaFree(na); if (ni) cudaFree(ni); return code; };
    if (cudaMalloc((void **)&np, cap * 3 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&nw, cap * 3 * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&na, cap * 32 * sizeof(float)) != cudaSuccess || cudaMalloc((void **)&ni, cap * sizeof(uint64_t)) != cudaSuccess) {
        ctx->err = "pv_allgather_photons: out of device memory for the gathered planes"; cudaGetLastError(); return fail(PV_ENOMEM);
    }
    // 4. the collective: four all-gathers in one group
    cudaEvent_t e0 = ctx->tev[0], e1 = ctx->tev[1];
    cudaEventRecord(e0, ctx->stream);
    ncclResult_t r0 = api->GroupStart();
    ncclResult_t r1 = api->AllGather(ctx->d_pos, np, mx * 3, ncclFloat32, comm, ctx->stream);
    ncclResult_t r2 = api->AllGather(ctx->d_wi, nw, mx * 3, ncclFloat32, comm, ctx->stream);
    ncclResult_t r3 = api->AllGather(ctx->d_alpha, na, mx * 32, ncclFloat32, comm, ctx->stream);
    ncclResult_t r4 = api->AllGather(ctx->d_ids, ni, mx, ncclUint64, comm, ctx->stream);
    ncclResult_t r5 = api->GroupEnd();
    cudaEventRecord(e1, ctx->stream);
    for (ncclResult_t r : {r0, r1, r2, r3, r4, r5})
        if (r != ncclSuccess) { ctx->err = std::string("pv_allgather_photons: ") + api->GetErrorString(r); return fail(PV_ECUDA); }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = std::string("pv_allgather_photons: ") + cudaGetErrorString(cudaGetLastError()); return fail(PV_ECUDA); }
    rc = comm_async_check(ctx, api, "pv_allgather_photons"); if (rc) return fail(rc);
    if (collective_ms) cudaEventElapsedTime(collective_ms, e0, e1);
    // 5. the gathered planes become the context's photon set; one sort by id orders it and drops the slack
    cudaFree(ctx->d_pos); cudaFree(ctx->d_wi); cudaFree(ctx->d_alpha); cudaFree(ctx->d_ids);
    ctx->d_pos = np; ctx->d_wi = nw; ctx->d_alpha = na; ctx->d_ids = ni; ctx->cap_photons = cap; ctx->n_photons = cap;
    rc = pvi_sort_photons_by_id(ctx); if (rc) return rc;
    if (ctx->n_photons != total) { ctx->err = "pv_allgather_photons: photon count after the merge does not match the ranks' counts"; return PV_ECUDA; }
    return PV_OK;
}

// One process, several devices: ctxs[src]'s photon set (planes in deposit order, ids included) to every other context.
int pvi_broadcast_photons(pv_ctx **ctxs, int n, int src, float *collective_ms) {
    if (collective_ms) *collective_ms = 0.f;
    pv_ctx *s = ctxs[src];
    const NcclApi *api = nccl_api(&s->err);
    if (!api) return PV_ESTATE;
    for (int i = 0; i < n; ++i) if (!ctxs[i]->comm || ctxs[i]->comm_world != n || ctxs[i]->comm_rank != i) {
        s->err = "pv_broadcast_photons: the contexts do not share a communicator (pv_comm_init_all first)"; return PV_ESTATE;
    }
    const uint64_t cnt = s->n_photons;
    for (int i = 0; i < n; ++i) {
        pv_ctx *c = ctxs[i];
        cudaSetDevice(c->device);
        if (i != src) { c->built = false; c->n_photons = 0; int rc = pvi_reserve_photons(c, cnt); if (rc) { s->err = c->err; return rc; } }     // (the source keeps its map)
    }
    if (cnt == 0) { for (int i = 0; i < n; ++i) ctxs[i]->n_photons = 0; return PV_OK; }
    cudaSetDevice(s->device);
    cudaEventRecord(s->tev[0], s->stream);
    ncclResult_t bad = api->GroupStart();
    for (int i = 0; i < n && bad == ncclSuccess; ++i) {
        pv_ctx *c = ctxs[i];
        ncclComm_t comm = (ncclComm_t)c->comm;
        bad = api->Broadcast(s->d_pos, c->d_pos, cnt * 3, ncclFloat32, src, comm, c->stream);
        if (bad == ncclSuccess) bad = api->Broadcast(s->d_wi, c->d_wi, cnt * 3, ncclFloat32, src, comm, c->stream);
        if (bad == ncclSuccess) bad = api->Broadcast(s->d_alpha, c->d_alpha, cnt * 32, ncclFloat32, src, comm, c->stream);
        if (bad == ncclSuccess) bad = api->Broadcast(s->d_ids, c->d_ids, cnt, ncclUint64, src, comm, c->stream);
    }
    ncclResult_t ge = api->GroupEnd();
    if (bad == ncclSuccess) bad = ge;
    if (bad != ncclSuccess) { s->err = std::string("pv_broadcast_photons: ") + api->GetErrorString(bad); return PV_ECUDA; }
    cudaSetDevice(s->device);
    cudaEventRecord(s->tev[1], s->stream);
    for (int i = 0; i < n; ++i) {
        pv_ctx *c = ctxs[i];
        cudaSetDevice(c->device);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { s->err = std::string("pv_broadcast_photons: ") + cudaGetErrorString(cudaGetLastError()); return PV_ECUDA; }
        int rc = comm_async_check(c, api, "pv_broadcast_photons"); if (rc) { s->err = c->err; return rc; }
        c->n_photons = cnt;
        for (int k = 0; k < 4; ++k) c->map_paths[k] = s->map_paths[k];
    }
    cudaSetDevice(s->device);
    if (collective_ms) cudaEventElapsedTime(collective_ms, s->tev[0], s->tev[1]);
    return PV_OK;
}
