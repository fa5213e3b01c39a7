// pv_shoot.cu -- photon shooting (K1-K3), host side: PhotonShooter::Preprocess / PhotonShootingTask::Run's outer loop
// (core/photonshooter.cpp:232-357, 457-503).  Light paths are dealt in the reference's blocks of 4096; a deposit of global block b
// is divided by nshot = 4096*b (Q4), independent of the number of ranks.  A WAVE of blocks is traced by the wavefront kernels of
// pv_wavefront.cu (followPhoton itself, :47-229); the host sizes the photon buffer, replays a wave whose buffer or continuation-stack
// pool turned out too small (waves are deterministic), replays the reference's per-block bookkeeping of the done flags over the
// per-block deposit counts, and finally orders the photons by id = (class, path, deposit ordinal) so that the set depends only on
// (scene, seed, target).
#include <algorithm>
#include <type_traits>
#include <vector>
#include "pv_ctx.h"

#include <cstdlib>
#include <chrono>
#include "pv_shoot.cuh"
// PV_TIMING=1: wall time of the host-side phases of shooting on stderr (tuning aid)
namespace {
struct PhaseTimer {
    const char *what; bool on; std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(const char *w) : what(w), on(getenv("PV_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char *phase) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[pv timing] %s: %s %.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}

// ---------------------------------------------------------------- host: MT19937 only to reproduce task 0's Halton tables
static void halton_tables_task0(uint32_t perm[41]) {
    // RNG rng(31 * taskNum) with taskNum == 0, then PermutedHalton(6, rng): core/rng.cpp:43-107, montecarlo.cpp:380-397
    uint32_t mt[624]; int mti;
    mt[0] = 0u;
    for (mti = 1; mti < 624; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
    auto next = [&]() -> uint32_t {
        static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
        uint32_t y;
        if (mti >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u]; }
            for (; kk < 623; kk++) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u]; }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu); mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    };
    const uint32_t bases[6] = {2, 3, 5, 7, 11, 13};
    uint32_t *p = perm;
    for (int d = 0; d < 6; ++d) {
        uint32_t b = bases[d];
        for (uint32_t i = 0; i < b; ++i) p[i] = i;
        for (uint32_t i = 0; i < b; ++i) { uint32_t other = i + (next() % (b - i)); std::swap(p[i], p[other]); }
        p += b;
    }
}

// One wave of blocks.  surf_flags < 0: the volume-only kernel, counts[n_blocks].  Otherwise the all-maps kernel with the
// done flags SF_* held constant over the wave, counts[PC_COUNT + 1][n_blocks] (last row: first-hit scatter events, :104).
static int shoot_wave(pv_ctx *ctx, uint64_t first_block, uint32_t n_blocks, const pv_shoot_params *prm, int surf_flags, uint32_t *counts,
                      pv_shoot_stats *stats) {
    const bool surf = surf_flags >= 0;
    const uint32_t n_cls = surf ? PC_COUNT + 1 : 1;
    if (!ctx->has_scene || ctx->hscene.med.type == PV_MEDIUM_NONE || ctx->hscene.n_lights == 0) {
        ctx->err = "pv_shoot: scene needs a medium and at least one light"; return PV_ESTATE;
    }
    if (first_block < 1 || prm->world < 1 || prm->rank >= prm->world) { ctx->err = "pv_shoot: bad block / rank arguments"; return PV_EINVAL; }
    if (!(prm->stepsize > 0.f) || !(prm->integrator_stepsize > 0.f)) { ctx->err = "pv_shoot: step sizes must be > 0"; return PV_EINVAL; }
    if (first_block == 1 && !surf) { ctx->n_photons = 0; ctx->built = false; }
    if (n_blocks == 0) return PV_OK;
    // blocks of this rank inside the wave: b with (b - 1) % world == rank
    uint64_t b_start = first_block + ((prm->rank + prm->world - ((first_block - 1) % prm->world)) % prm->world);
    uint64_t b_end = first_block + n_blocks;              // exclusive
    uint32_t n_local = b_start < b_end ? (uint32_t)((b_end - b_start + prm->world - 1) / prm->world) : 0;
    memset(counts, 0, sizeof(uint32_t) * n_blocks * n_cls);
    if (n_local == 0) return PV_OK;

    int rc = pv_ensure(ctx, &ctx->io2, &ctx->io2_bytes, sizeof(uint32_t) * (size_t)n_blocks * n_cls + 64); if (rc) return rc;
    uint32_t *d_counts = (uint32_t *)ctx->io2;
    unsigned long long *d_nout = ctx->d_counters + 1, *d_work = ctx->d_counters + 2, *d_stats = ctx->d_counters + 8;
    // first guess of the capacity on top of what is already stored: the deposits per path seen in this context's earlier waves
    // (with a margin), else 8% (volume only) / 150% (all maps); never more than half of the free device memory -- a wave that
    // overflows its buffer is replayed with the exact size anyway
    double &yield = ctx->shoot_yield[surf ? 1 : 0];
    const double per_path = yield > 0. ? yield * 1.25 : (surf ? 1.5 : 0.08);
    uint64_t want_cap = ctx->n_photons + (uint64_t)((double)n_local * SH_BLOCK * per_path) + 65536;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const uint64_t room = ctx->cap_photons + (uint64_t)(free_b / 2) / 160;         // 160 B per photon over the four planes
            if (want_cap > room && room > ctx->n_photons + 65536) want_cap = room;
        }
    }
    PhaseTimer tm("shoot_wave");
    for (int attempt = 0; attempt < 3; ++attempt) {
        rc = pvi_reserve_photons(ctx, want_cap); if (rc) return rc;
        tm.lap("reserve");
        ShootArgs a;
        a.sc = ctx->dscene; a.b_start = b_start; a.n_local_blocks = n_local; a.world = prm->world; a.first_block = first_block;
        a.stepsize = prm->stepsize; a.istep4 = 4.f * prm->integrator_stepsize; a.max_depth = prm->max_photon_depth;
        a.k0 = (uint32_t)prm->seed; a.k1 = (uint32_t)(prm->seed >> 32);
        halton_tables_task0(a.perm);
        a.pos = ctx->d_pos; a.wi = ctx->d_wi; a.alpha32 = ctx->d_alpha; a.ids = ctx->d_ids;
        a.n_out = d_nout; a.cap = ctx->cap_photons; a.block_counts = d_counts; a.work = d_work; a.stats = d_stats;
        a.wave_blocks = n_blocks; a.flags = surf ? (uint32_t)surf_flags : 0u;
        unsigned long long init_n = ctx->n_photons, zero = 0;
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_nout, &init_n, sizeof(init_n), cudaMemcpyHostToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(d_work, &zero, sizeof(zero), cudaMemcpyHostToDevice, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemsetAsync(d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * n_blocks * n_cls, ctx->stream));
        const int kind = (ctx->hscene.n_spheres != 0 ? 1 : 0) | (ctx->hscene.med.type == PV_MEDIUM_EXPONENTIAL ? 2 : 0);
        PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        bool replay = false;
        rc = pvi_wavefront_run(ctx, a, surf, kind, &replay); if (rc) return rc;
        PV_CUDA_CHECK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaGetLastError());
        unsigned long long h_nout = 0, h_stats[8];
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(&h_nout, d_nout, sizeof(h_nout), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaMemcpyAsync(counts, d_counts, sizeof(uint32_t) * n_blocks * n_cls, cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        tm.lap("kernels + readback");
        if (replay) { --attempt; continue; }               // (the pool is at most quadrupled log4(P) times)
        if (h_nout > ctx->cap_photons) { want_cap = h_nout + 65536; continue; }      // too small: grow and replay the (deterministic) wave
        float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        yield = std::max(yield, (double)(h_nout - ctx->n_photons) / ((double)n_local * SH_BLOCK));
        ctx->n_photons = h_nout;
        if (stats) {
            stats->nodes_visited += h_stats[0]; stats->tri_tests += h_stats[1]; stats->density_samples += h_stats[2];
            stats->segments += h_stats[3]; stats->stack_overflows += h_stats[4]; stats->paths_local += h_stats[5];
            stats->seconds += ms * 1e-3;
        }
        return PV_OK;
    }
    ctx->err = "pv_shoot: could not size the photon buffer";
    return PV_ENOMEM;
}
int pvi_shoot_blocks(pv_ctx *ctx, uint64_t first_block, uint32_t n_blocks, const pv_shoot_params *prm, uint32_t *counts, pv_shoot_stats *stats) {
    return shoot_wave(ctx, first_block, n_blocks, prm, -1, counts, stats);
}

__global__ void id_keys_kernel(const uint64_t *__restrict__ ids, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = ids[i]; vals[i] = (uint32_t)i; }
}
int pvi_reserve_set(pv_ctx *ctx, PhotonSet *s, uint64_t n) {
    if (n <= s->cap) return PV_OK;
    pvi_free_set(s);
    uint64_t cap = std::max<uint64_t>(n, 1024);
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->pos, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->wi, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->alpha, cap * 32 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&s->ids, cap * sizeof(uint64_t)));
    s->cap = cap;
    return PV_OK;
}
void pvi_free_set(PhotonSet *s) {
    if (s->pos) cudaFree(s->pos);
    if (s->wi) cudaFree(s->wi);
    if (s->alpha) cudaFree(s->alpha);
    if (s->ids) cudaFree(s->ids);
    *s = PhotonSet();
}
__global__ void permute_photons_kernel(const uint32_t *__restrict__ order, uint64_t n, const float *__restrict__ pos, const float *__restrict__ wi,
                                       const float *__restrict__ alpha, const uint64_t *__restrict__ ids, float *__restrict__ pos_o,
                                       float *__restrict__ wi_o, float *__restrict__ alpha_o, uint64_t *__restrict__ ids_o) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t j = t >> 3; uint32_t sub = (uint32_t)(t & 7);
    if (j >= n) return;
    uint32_t src = order[j];
    *(reinterpret_cast<float4 *>(alpha_o + j * 32) + sub) = *(reinterpret_cast<const float4 *>(alpha + (uint64_t)src * 32) + sub);
    if (sub < 3) { pos_o[3 * j + sub] = pos[3 * (uint64_t)src + sub]; wi_o[3 * j + sub] = wi[3 * (uint64_t)src + sub]; }
    if (sub == 3) ids_o[j] = ids[src];
}

// Drop photons of blocks > last_block and order the rest by id = (path index << 16 | deposit ordinal): the photon
// set and its order then depend only on (scene, seed, target), not on thread scheduling or the number of ranks.
// split: the ids carry a class in their top bits (pv_shoot_maps); class 0 stays the context's photon set, the others
// move to ctx->surf[class - 1].
static int shoot_finish(pv_ctx *ctx, uint64_t last_block, bool split, bool bound_always = false) {
    uint64_t n = ctx->n_photons;
    ctx->built = false;
    if (split) for (int c = 0; c < 4; ++c) ctx->surf[c].n = 0;
    if (n == 0) return PV_OK;
    if (n > 0xFFFFFFF0ull) { ctx->err = "too many photons"; return PV_EINVAL; }
    size_t need = n * (2 * sizeof(uint64_t) + 2 * sizeof(uint32_t)) + 256;
    int rc = pv_ensure(ctx, &ctx->scratch, &ctx->scratch_bytes, need); if (rc) return rc;
    uint64_t *keys = (uint64_t *)ctx->scratch, *keys_tmp = keys + n;
    uint32_t *vals = (uint32_t *)(keys_tmp + n), *vals_tmp = vals + n;
    PhaseTimer tm("shoot_finish");
    id_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_ids, n, keys, vals);
    PV_CUDA_CHECK(ctx, cudaGetLastError());
    // highest path index bounds the key width
    uint64_t max_path = last_block ? last_block * SH_BLOCK : ~0ull >> 16;
    int bits = 16; while (bits < 64 && (max_path >> (bits - 16)) != 0) ++bits;
    bits = std::min(64, bits + 1);
    // photons of later blocks may still be present: they sort to the end because their path index is larger
    uint64_t *skeys; uint32_t *svals;
    rc = pvi_sort_pairs_u64(ctx, keys, vals, keys_tmp, vals_tmp, n, 64, &skeys, &svals); if (rc) return rc;
    (void)bits;
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));      // the probes below are blocking copies on another stream
    tm.lap("sort by id");
    // count survivors: ids <= (last_block * 4096) << 16 | 0xffff
    auto upper_bound = [&](uint64_t limit, uint64_t *out) -> int {  // first position with key > limit (few D2H probes)
        uint64_t probe = 0, lo = 0, hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            PV_CUDA_CHECK(ctx, cudaMemcpy(&probe, skeys + mid, sizeof(uint64_t), cudaMemcpyDeviceToHost));
            if (probe <= limit) lo = mid + 1; else hi = mid;
        }
        *out = lo;
        return PV_OK;
    };
    const uint64_t path_limit = last_block ? (((last_block * SH_BLOCK) << 16) | 0xffffull) : ((1ull << 60) - 1);
    uint64_t keep = n;
    if (last_block || split || bound_always) { rc = upper_bound(path_limit, &keep); if (rc) return rc; }
    if (split) {
        for (uint32_t c = 1; c < PC_COUNT; ++c) {
            uint64_t lo = 0, hi = 0;
            rc = upper_bound(((uint64_t)c << 60) - 1, &lo); if (rc) return rc;
            rc = upper_bound(((uint64_t)c << 60) | path_limit, &hi); if (rc) return rc;
            PhotonSet &ps = ctx->surf[c - 1];
            rc = pvi_reserve_set(ctx, &ps, hi - lo); if (rc) return rc;
            ps.n = hi - lo;
            if (ps.n) {
                permute_photons_kernel<<<(unsigned)((ps.n * 8 + 255) / 256), 256, 0, ctx->stream>>>(svals + lo, ps.n, ctx->d_pos, ctx->d_wi, ctx->d_alpha,
                                                                                                  ctx->d_ids, ps.pos, ps.wi, ps.alpha, ps.ids);
                PV_CUDA_CHECK(ctx, cudaGetLastError());
            }
        }
    }
    tm.lap("bounds");
    float *np, *nw, *na; uint64_t *ni;
    uint64_t cap = std::max<uint64_t>(keep, 1024);
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&np, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&nw, cap * 3 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&na, cap * 32 * sizeof(float)));
    PV_CUDA_CHECK(ctx, cudaMalloc((void **)&ni, cap * sizeof(uint64_t)));
    tm.lap("malloc");
    if (keep) {
        permute_photons_kernel<<<(unsigned)((keep * 8 + 255) / 256), 256, 0, ctx->stream>>>(svals, keep, ctx->d_pos, ctx->d_wi, ctx->d_alpha, ctx->d_ids,
                                                                                         np, nw, na, ni);
        PV_CUDA_CHECK(ctx, cudaGetLastError());
    }
    PV_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    tm.lap("permute");
    cudaFree(ctx->d_pos); cudaFree(ctx->d_wi); cudaFree(ctx->d_alpha); cudaFree(ctx->d_ids);
    ctx->d_pos = np; ctx->d_wi = nw; ctx->d_alpha = na; ctx->d_ids = ni; ctx->cap_photons = cap; ctx->n_photons = keep;
    tm.lap("free");
    return PV_OK;
}
int pvi_shoot_finish(pv_ctx *ctx, uint64_t last_block) { return shoot_finish(ctx, last_block, false); }
// order the context's photon set by id and drop the records whose id is not a volume-photon id (class bits set: the slack of
// pv_allgather_photons carries id = ~0)
int pvi_sort_photons_by_id(pv_ctx *ctx) { return shoot_finish(ctx, 0, false, true); }

// Single-rank driver == PhotonShootingTask::Run's outer loop (photonshooter.cpp:245-356): waves of blocks until the
// running photon count reaches the target at some block M; give-up rule of :285-299.
int pvi_shoot(pv_ctx *ctx, uint64_t n_wanted, const pv_shoot_params *prm_in, pv_shoot_stats *stats) {
    pv_shoot_params prm = *prm_in;
    pv_shoot_stats st; memset(&st, 0, sizeof(st));
    if (prm.world != 1 || prm.rank != 0) { ctx->err = "pv_shoot: use pv_shoot_blocks/pv_shoot_finish when world > 1"; return PV_EINVAL; }
    uint64_t max_paths = prm.max_paths ? prm.max_paths : ((uint64_t)1 << 40);
    uint64_t block = 0, total = 0, last = 0;
    uint32_t wave = 64;
    std::vector<uint32_t> counts;
    bool done = false; int rc = PV_OK;
    if (n_wanted == 0) { ctx->n_photons = 0; ctx->built = false; if (stats) *stats = st; return PV_OK; }
    while (!done) {
        counts.assign(wave, 0);
        rc = pvi_shoot_blocks(ctx, block + 1, wave, &prm, counts.data(), &st); if (rc) return rc;
        for (uint32_t i = 0; i < wave; ++i) {
            uint64_t nshot_before = block * SH_BLOCK;
            // "Unable to store enough photons.  Giving up." (photonshooter.cpp:285-299, unsuccessful() :37-39)
            if (nshot_before > 500000 && total < n_wanted && (total == 0 || total < SH_BLOCK / 1024)) {
                ctx->n_photons = 0; ctx->err = "Unable to store enough photons.  Giving up."; rc = PV_ENOPHOTONS; done = true; break;
            }
            block++; total += counts[i];
            if (total >= n_wanted || block * SH_BLOCK >= max_paths) { last = block; done = true; break; }
        }
        if (!done) {
            // size the next wave from the observed yield, aiming a little past the target
            double per_block = std::max(1e-3, (double)total / (double)block);
            double remaining = (double)(n_wanted - total) / per_block;
            wave = (uint32_t)std::min<double>(std::max<double>(remaining * 1.03 + 8, 64), 262144);
            uint64_t left = (max_paths / SH_BLOCK > block) ? (max_paths / SH_BLOCK - block) : 1;
            wave = (uint32_t)std::min<uint64_t>(wave, left);
        }
    }
    if (rc == PV_OK) rc = pvi_shoot_finish(ctx, last);
    st.paths = last * SH_BLOCK; st.blocks = last; st.photons_local = ctx->n_photons;
    if (stats) *stats = st;
    return rc;
}

// All photon maps in one pass == PhotonShootingTask::Run's outer loop for ONE task (photonshooter.cpp:245-356).  The
// reference's done flags change how later paths behave (no more scattering in the medium once the volume map is full,
// :96; diffuse bounces end the path once the indirect map is full, :218), and they change at block ends.  A wave of
// blocks is traced with the flags held constant; the host then replays the reference's per-block bookkeeping over the
// per-class counts, and if a flag flips at block M inside the wave, the wave is rolled back and traced again up to M
// (it is deterministic), so that every later block sees the new flags.  Waves aim just short of the next expected flip.
// Several ranks: every rank traces the blocks b with (b - 1) % world == rank of each wave, the per-class, per-block counts are
// summed over ranks through `allreduce`, and every rank then replays the SAME bookkeeping on the same numbers -- flags, flips,
// roll-backs and the last block come out identical everywhere without any other communication.
int pvi_shoot_maps(pv_ctx *ctx, const pv_maps_params *mp, const pv_shoot_params *prm_in, pv_allreduce_u32_fn allreduce, void *user,
                   pv_maps_stats *out) {
    pv_shoot_params prm = *prm_in;
    pv_maps_stats ms; memset(&ms, 0, sizeof(ms));
    if (prm.world < 1 || prm.rank >= prm.world) { ctx->err = "pv_shoot_maps: bad rank / world"; return PV_EINVAL; }
    if (prm.world > 1 && !allreduce) { ctx->err = "pv_shoot_maps: world > 1 needs pv_shoot_maps_ranks with an all-reduce callback"; return PV_EINVAL; }
    const uint64_t wanted[3] = {mp->n_volume_wanted, mp->n_caustic_wanted, mp->n_indirect_wanted};     // by class id
    bool done[3] = {wanted[0] == 0, wanted[1] == 0, wanted[2] == 0};
    ctx->n_photons = 0; ctx->built = false; ctx->rad_valid = false;
    for (int c = 0; c < 4; ++c) { ctx->surf[c].n = 0; ctx->map_paths[c] = 0; }
    if (done[0] && done[1] && done[2]) { if (out) *out = ms; return PV_OK; }
    const uint64_t max_paths = prm.max_paths ? prm.max_paths : ((uint64_t)1 << 40);
    uint64_t block = 0, tot[PC_COUNT] = {0, 0, 0, 0, 0}, first_hits = 0;
    uint64_t paths[4] = {0, 0, 0, 0};                       // caustic, indirect, direct, volume
    uint32_t wave = 64;
    std::vector<uint32_t> counts;
    bool finished = false, aborted = false;
    int rc = PV_OK;
    auto unsuccessful = [](uint64_t needed, uint64_t found) { return found < needed && (found == 0 || found < SH_BLOCK / 1024); };
    while (!finished) {
        const int flags = (done[1] ? 0 : SF_WANT_CAUSTIC) | (done[2] ? 0 : SF_WANT_INDIRECT) | (done[0] ? SF_VOLUME_DONE : 0) |
                          (mp->final_gather ? SF_FINAL_GATHER : 0);
        const uint64_t n_before = ctx->n_photons, first = block + 1;
        counts.assign((size_t)wave * (PC_COUNT + 1), 0);
        rc = shoot_wave(ctx, first, wave, &prm, flags, counts.data(), &ms.shoot); if (rc) return rc;
        if (prm.world > 1 && allreduce(counts.data(), counts.size(), user) != 0) { ctx->err = "pv_shoot_maps: the all-reduce callback failed"; return PV_EINVAL; }
        bool flip = false;
        uint32_t used = 0;
        for (uint32_t i = 0; i < wave; ++i) {
            // "Unable to store enough photons.  Giving up." over all three wanted counts (:285-299)
            if (block * SH_BLOCK > 500000 && (unsuccessful(wanted[1], tot[1]) || unsuccessful(wanted[2], tot[2]) || unsuccessful(wanted[0], tot[0]))) {
                aborted = true; finished = true; break;
            }
            block++; used++;
            if (!done[2]) {                                  // :303-318
                paths[1] += SH_BLOCK; tot[2] += counts[(size_t)PC_INDIRECT * wave + i];
                paths[2] += SH_BLOCK; tot[3] += counts[(size_t)PC_DIRECT * wave + i];
                if (tot[2] >= wanted[2]) { done[2] = true; flip = true; }
            }
            if (!done[1]) {                                  // :320-328
                paths[0] += SH_BLOCK; tot[1] += counts[(size_t)PC_CAUSTIC * wave + i];
                if (tot[1] >= wanted[1]) { done[1] = true; flip = true; }
            }
            if (!done[0]) {                                  // :330-341
                paths[3] += SH_BLOCK; tot[0] += counts[(size_t)PC_VOLUME * wave + i];
                if (tot[0] >= wanted[0]) { done[0] = true; flip = true; }
            }
            tot[4] += counts[(size_t)PC_RADIANCE * wave + i];
            first_hits += counts[(size_t)PC_COUNT * wave + i];
            if ((done[0] && done[1] && done[2]) || block * SH_BLOCK >= max_paths) { finished = true; break; }
            if (flip) break;
        }
        if (aborted) break;
        if (flip && !finished && used < wave) {
            // roll the wave back and trace it again up to the block of the flip, with the flags it started with
            ctx->n_photons = n_before;
            counts.assign((size_t)used * (PC_COUNT + 1), 0);
            rc = shoot_wave(ctx, first, used, &prm, flags, counts.data(), &ms.shoot); if (rc) return rc;
            ms.replayed_blocks += used;                      // (the counts of the replay are the ones already booked: no second all-reduce)
        }
        if (!finished) {
            // next wave: 97% of the way to the nearest expected flip, from the yields seen so far
            double nearest = 1e30;
            for (int c = 0; c < 3; ++c)
                if (!done[c]) {
                    const double per_block = (double)tot[c] / (double)block;
                    nearest = std::min(nearest, per_block > 0 ? (double)(wanted[c] - tot[c]) / per_block : 4.0 * (double)block);
                }
            wave = (uint32_t)std::min<double>(std::max<double>(nearest * 0.97, 16), 65536);
            const uint64_t left = max_paths / SH_BLOCK > block ? max_paths / SH_BLOCK - block : 1;
            wave = (uint32_t)std::min<uint64_t>(wave, left);
        }
    }
    if (aborted) {
        ctx->n_photons = 0;
        for (int c = 0; c < 4; ++c) ctx->surf[c].n = 0;
        ctx->err = "Unable to store enough photons.  Giving up.";
        rc = PV_ENOPHOTONS;
    } else rc = shoot_finish(ctx, block, true);
    ms.nshot = block * SH_BLOCK; ms.blocks = block;
    ms.n_caustic_paths = paths[0]; ms.n_indirect_paths = paths[1]; ms.n_direct_paths = paths[2]; ms.n_volume_paths = paths[3] + first_hits;
    ms.shoot.paths = ms.nshot; ms.shoot.blocks = block; ms.shoot.photons_local = ctx->n_photons;
    ms.n[0] = ctx->n_photons;
    for (int c = 0; c < 4; ++c) { ms.n[c + 1] = ctx->surf[c].n; ctx->map_paths[c] = c == 3 ? ms.n_volume_paths : paths[c]; }
    if (out) *out = ms;
    return rc;
}
