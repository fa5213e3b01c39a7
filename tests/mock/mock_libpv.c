/* mock_libpv.c -- TEST DOUBLE of csrc/libpv.so for CPU tests of the drop-in's HOST logic (tests/test_adapter_host.py).
 *
 * Not a CPU implementation of anything: the "radiance" it returns is a hash of the ray's global stream index, which is exactly
 * what makes it useful -- an image rendered through it changes if the adapter hands a ray to the device under another index,
 * drops a ray, or stitches device results back in the wrong place.  Every call is appended to $MOCK_PV_LOG.
 * Photon maps it "shoots" are synthetic points on the floor of the test scenes' box.
 * Exports the entry points host/pv_pbrt_adapter.cpp binds (include/pv.h). */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <fcntl.h>
#include "pv.h"

struct pv_ctx { int device; uint64_t n_photons; double photon_sum; uint64_t n_map[5]; int selected; };

static void logf_(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
#include <stdarg.h>
static void logf_(const char *fmt, ...) {
    const char *fn = getenv("MOCK_PV_LOG");
    if (!fn) return;
    char buf[512];
    va_list ap; va_start(ap, fmt); int n = vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    int fd = open(fn, O_WRONLY | O_APPEND | O_CREAT, 0644);
    if (fd >= 0) { if (write(fd, buf, (size_t)n) < 0) {} close(fd); }          /* one O_APPEND write per line: atomic across threads */
}

int pv_version(void) { return 100; }
const char *pv_last_error(pv_ctx *ctx) { (void)ctx; return "mock"; }
int pv_create(pv_ctx **out, int device) {
    *out = (pv_ctx *)calloc(1, sizeof(pv_ctx)); (*out)->device = device;
    logf_("create dev=%d\n", device);
    return PV_OK;
}
void pv_destroy(pv_ctx *ctx) { if (ctx) logf_("destroy dev=%d\n", ctx->device); free(ctx); }
int pv_set_scene(pv_ctx *ctx, const pv_scene_desc *s) {
    /* order= : per position of the primitive array, a fingerprint of the triangle there (so a test sees a permutation) */
    char buf[4096]; int o = 0; buf[0] = 0;
    for (uint32_t i = 0; i < s->n_prims && i < 24 && o < 400; ++i) {
        double f = 0.; for (int k = 0; k < 9; ++k) f += (k + 1) * (double)s->tri_verts[9 * i + k];
        o += snprintf(buf + o, sizeof(buf) - o, "%s%ld", i ? "," : "", (long)(1000. * f) * 8 + (long)(1000.f * s->materials[s->prim_material[i]].kd[29]) % 8);
    }
    logf_("set_scene dev=%d prims=%u lights=%u nodes=%u order:%s\n", ctx->device, s->n_prims, s->n_lights, s->n_nodes, buf);
    return PV_OK;
}
/* the double's "BVH": one leaf holding everything (scenes of the host tests have a dozen triangles); MOCK_PV_BVH_REVERSE=1 hands the
 * primitives back in reverse order */
int pv_build_bvh(pv_ctx *ctx, const float *pb, uint32_t n, uint32_t max_prims, pv_bvh_node *nodes, uint32_t cap, uint32_t *n_nodes,
                 uint32_t *order, float *ms) {
    const char *rev = getenv("MOCK_PV_BVH_REVERSE");
    logf_("build_bvh dev=%d n=%u maxprims=%u cap=%u\n", ctx->device, n, max_prims, cap);
    if (n > 255 || cap < 1) return PV_EINVAL;
    memset(&nodes[0], 0, sizeof(nodes[0]));
    for (int k = 0; k < 3; ++k) { nodes[0].bounds[k] = 1e30f; nodes[0].bounds[3 + k] = -1e30f; }
    for (uint32_t i = 0; i < n; ++i) {
        order[i] = (rev && rev[0] == '1') ? n - 1 - i : i;
        for (int k = 0; k < 3; ++k) {
            if (pb[6 * i + k] < nodes[0].bounds[k]) nodes[0].bounds[k] = pb[6 * i + k];
            if (pb[6 * i + 3 + k] > nodes[0].bounds[3 + k]) nodes[0].bounds[3 + k] = pb[6 * i + 3 + k];
        }
    }
    nodes[0].offset = 0; nodes[0].n_primitives = (uint8_t)n;
    *n_nodes = n ? 1 : 0; if (ms) *ms = 0.f;
    return PV_OK;
}
int pv_shoot(pv_ctx *ctx, uint64_t wanted, const pv_shoot_params *p, pv_shoot_stats *st) {
    (void)p; ctx->n_photons = wanted; if (st) { memset(st, 0, sizeof(*st)); st->paths = 4096; }
    logf_("shoot dev=%d n=%llu\n", ctx->device, (unsigned long long)wanted);
    return PV_OK;
}
/* every map "filled" to what was asked for; 300 radiance-photon sites when final gathering is on */
int pv_shoot_maps(pv_ctx *ctx, const pv_maps_params *mp, const pv_shoot_params *p, pv_maps_stats *st) {
    (void)p; ctx->n_photons = mp->n_volume_wanted;
    ctx->n_map[PV_MAP_VOLUME] = mp->n_volume_wanted; ctx->n_map[PV_MAP_CAUSTIC] = mp->n_caustic_wanted;
    ctx->n_map[PV_MAP_INDIRECT] = mp->n_indirect_wanted; ctx->n_map[PV_MAP_DIRECT] = 0; ctx->n_map[PV_MAP_RADIANCE] = mp->final_gather ? 300 : 0;
    if (st) {
        memset(st, 0, sizeof(*st));
        for (int k = 0; k < 5; ++k) st->n[k] = ctx->n_map[k];
        st->nshot = 4096; st->blocks = 1; st->n_caustic_paths = st->n_indirect_paths = st->n_direct_paths = st->n_volume_paths = 4096;
    }
    logf_("shoot_maps dev=%d volume=%llu caustic=%llu indirect=%llu fg=%d\n", ctx->device, (unsigned long long)mp->n_volume_wanted,
          (unsigned long long)mp->n_caustic_wanted, (unsigned long long)mp->n_indirect_wanted, (int)mp->final_gather);
    return PV_OK;
}
int pv_build(pv_ctx *ctx, float maxdist, uint32_t nused) { logf_("build dev=%d n=%llu maxdist=%g nused=%u\n", ctx->device, (unsigned long long)ctx->n_photons, maxdist, nused); return PV_OK; }
int pv_photon_count(pv_ctx *ctx, uint64_t *n) { *n = ctx->n_photons; return PV_OK; }
int pv_get_photons(pv_ctx *ctx, float *pos, float *wi, float *alpha, uint64_t *ids, uint64_t cap, uint64_t *n) {
    uint64_t m = cap < ctx->n_photons ? cap : ctx->n_photons;
    double sum = 0;
    for (uint64_t i = 0; i < m; ++i) {
        for (int k = 0; k < 3; ++k) { if (pos) { pos[3 * i + k] = (float)(i % 977) * 0.001f + k; sum += pos[3 * i + k]; } if (wi) wi[3 * i + k] = k == 2; }
        for (int b = 0; b < PV_NSPEC && alpha; ++b) { alpha[PV_NSPEC * i + b] = 1e-3f * (float)(b + 1); sum += alpha[PV_NSPEC * i + b]; }
        if (ids) ids[i] = i;
    }
    if (n) *n = m;
    logf_("get_photons dev=%d n=%llu sum=%.6f\n", ctx->device, (unsigned long long)m, sum);
    return PV_OK;
}
int pv_set_photons(pv_ctx *ctx, const float *pos, const float *wi, const float *alpha, uint64_t n) {
    (void)wi; double sum = 0;
    for (uint64_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3 && pos; ++k) sum += pos[3 * i + k];
        for (int b = 0; b < PV_NSPEC && alpha; ++b) sum += alpha[PV_NSPEC * i + b];
    }
    ctx->n_photons = n; ctx->photon_sum = sum;
    logf_("set_photons dev=%d n=%llu sum=%.6f\n", ctx->device, (unsigned long long)n, sum);
    return PV_OK;
}
/* the map's replication inside one process (csrc/pv_comm.cu): one communicator over the contexts, one broadcast */
static int g_comm_n = 0;
int pv_comm_init_all(pv_ctx **ctxs, int n) {
    char devs[128] = ""; for (int i = 0; i < n; ++i) { char t[16]; snprintf(t, sizeof t, "%s%d", i ? "," : "", ctxs[i]->device); strcat(devs, t); }
    g_comm_n = n; logf_("comm_init_all n=%d devs=%s\n", n, devs); return PV_OK;
}
int pv_comm_destroy(pv_ctx *ctx) { (void)ctx; return PV_OK; }
int pv_broadcast_photons(pv_ctx **ctxs, int n, int src, float *ms) {
    if (n != g_comm_n) return PV_ESTATE;
    for (int i = 0; i < n; ++i) { ctxs[i]->n_photons = ctxs[src]->n_photons; ctxs[i]->photon_sum = ctxs[src]->photon_sum; }
    if (ms) *ms = 0.f;
    logf_("broadcast_photons dev=%d n=%llu to=%d\n", ctxs[src]->device, (unsigned long long)ctxs[src]->n_photons, n - 1);
    return PV_OK;
}
static void fake_li(const char *what, pv_ctx *ctx, const pv_ray *rays, uint64_t n, const pv_gather_params *p, float *L, float *T) {
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t g = p->ray_index_base + i;
        uint32_t h = (uint32_t)(g * 2654435761u) ^ (uint32_t)(g >> 32);
        for (int b = 0; b < PV_NSPEC; ++b) {
            L[PV_NSPEC * i + b] = (float)((h >> (b % 16)) & 0xffu) / 255.f * (rays[i].d[2] > 0.f ? 1.f : .5f);
            T[PV_NSPEC * i + b] = 0.75f;
        }
    }
    logf_("%s dev=%d base=%llu n=%llu\n", what, ctx->device, (unsigned long long)p->ray_index_base, (unsigned long long)n);
}
int pv_gather(pv_ctx *ctx, const pv_ray *rays, uint64_t n, const pv_gather_params *p, float *L, float *T) { fake_li("gather", ctx, rays, n, p, L, T); return PV_OK; }
int pv_volume_li(pv_ctx *ctx, int integrator, const pv_ray *rays, uint64_t n, const pv_gather_params *p, float *L, float *T) {
    fake_li(integrator == PV_VOLINT_SINGLE ? "volume_li_single" : "volume_li_emission", ctx, rays, n, p, L, T); return PV_OK;
}
/* the indexed forms: every ray under its own stream index */
static void fake_li_indexed(const char *what, pv_ctx *ctx, const pv_ray *rays, const uint64_t *idx, uint64_t n, float *L, float *T) {
    for (uint64_t i = 0; i < n; ++i) {
        pv_gather_params p; memset(&p, 0, sizeof(p)); p.ray_index_base = idx[i];
        const char *keep = getenv("MOCK_PV_LOG"); (void)keep;
        uint64_t g = idx[i];
        uint32_t h = (uint32_t)(g * 2654435761u) ^ (uint32_t)(g >> 32);
        for (int b = 0; b < PV_NSPEC; ++b) {
            L[PV_NSPEC * i + b] = (float)((h >> (b % 16)) & 0xffu) / 255.f * (rays[i].d[2] > 0.f ? 1.f : .5f);
            T[PV_NSPEC * i + b] = 0.75f;
        }
    }
    logf_("%s dev=%d base=%llu n=%llu indexed=1\n", what, ctx->device, (unsigned long long)(n ? idx[0] : 0), (unsigned long long)n);
}
int pv_gather_indexed(pv_ctx *ctx, const pv_ray *rays, const uint64_t *idx, uint64_t n, const pv_gather_params *p, float *L, float *T) {
    (void)p; fake_li_indexed("gather", ctx, rays, idx, n, L, T); return PV_OK;
}
int pv_volume_li_indexed(pv_ctx *ctx, int integrator, const pv_ray *rays, const uint64_t *idx, uint64_t n, const pv_gather_params *p, float *L, float *T) {
    (void)p; fake_li_indexed(integrator == PV_VOLINT_SINGLE ? "volume_li_single" : "volume_li_emission", ctx, rays, idx, n, L, T); return PV_OK;
}
int pv_last_kernel_ms(pv_ctx *ctx, float *ms) { (void)ctx; *ms = 0.f; return PV_OK; }
/* surface-map entry points: present so the binary loads; the host-logic tests use scenes without surface photon maps */
/* surface photons / radiance-photon sites scattered over the floor of the Cornell box (y = -1), arriving from straight above */
int pv_get_map_photons(pv_ctx *c, int m, float *pos, float *wi, float *alpha, uint64_t *ids, uint64_t cap, uint64_t *n) {
    uint64_t k = cap < c->n_map[m] ? cap : c->n_map[m];
    for (uint64_t i = 0; i < k; ++i) {
        uint32_t h = (uint32_t)(i * 2654435761u + (uint32_t)m * 40503u);
        if (pos) { pos[3 * i] = (float)(h & 0xffffu) / 32768.f - 1.f; pos[3 * i + 1] = -1.f; pos[3 * i + 2] = (float)(h >> 16) / 32768.f - 1.f; }
        if (wi) { wi[3 * i] = 0.f; wi[3 * i + 1] = 1.f; wi[3 * i + 2] = 0.f; }
        for (int b = 0; b < PV_NSPEC && alpha; ++b) alpha[PV_NSPEC * i + b] = 0.002f;
        if (ids) ids[i] = ((uint64_t)m << 60) | i;
    }
    if (n) *n = k;
    logf_("get_map_photons dev=%d map=%d n=%llu\n", c->device, m, (unsigned long long)k);
    return PV_OK;
}
int pv_set_map_photons(pv_ctx *c, int m, const float *a, const float *b, const float *d, uint64_t n) {
    (void)a; (void)b; (void)d; c->n_map[m] = n; logf_("set_map_photons dev=%d map=%d n=%llu\n", c->device, m, (unsigned long long)n); return PV_OK;
}
int pv_radiance_photons(pv_ctx *c, uint32_t k, float r2, const uint64_t *pc, float *Lo, uint64_t cap, uint64_t *n) {
    (void)k; (void)r2; (void)pc;
    uint64_t m = cap < c->n_map[PV_MAP_RADIANCE] ? cap : c->n_map[PV_MAP_RADIANCE];
    for (uint64_t i = 0; i < m * PV_NSPEC && Lo; ++i) Lo[i] = 0.25f;
    if (n) *n = m;
    logf_("radiance_photons dev=%d n=%llu\n", c->device, (unsigned long long)m);
    return PV_OK;
}
int pv_set_radiance_lo(pv_ctx *c, const float *Lo, uint64_t n) { (void)c; (void)Lo; (void)n; return PV_OK; }
int pv_select_map(pv_ctx *c, int m, float r, uint32_t k) { (void)r; (void)k; c->selected = m; logf_("select_map dev=%d map=%d\n", c->device, m); return PV_OK; }
/* shadow rays: a ray is "occluded" iff a hash of its origin says so (1 in 4); "transmittance" = a function of its length */
int pv_occluded(pv_ctx *c, const pv_ray *r, uint64_t n, uint8_t *hit) {
    for (uint64_t i = 0; i < n; ++i) hit[i] = ((int)(fabsf(r[i].o[0] * 13.f + r[i].o[1] * 5.f + r[i].o[2] * 3.f) * 8.f) % 4) == 0;
    logf_("occluded dev=%d n=%llu\n", c->device, (unsigned long long)n);
    return PV_OK;
}
int pv_transmittance(pv_ctx *c, const pv_ray *r, uint64_t n, float step, const float *u, float *T) {
    int with_u = u != NULL;
    for (uint64_t i = 0; i < n; ++i) {
        float len = r[i].maxt < 1e30f ? r[i].maxt : 1.f;
        for (int b = 0; b < PV_NSPEC; ++b) T[PV_NSPEC * i + b] = 1.f / (1.f + 0.1f * len);
        if (u && !(u[i] >= 0.f && u[i] < 1.f)) with_u = 2;
    }
    logf_("transmittance dev=%d n=%llu step1000=%d u=%d\n", c->device, (unsigned long long)n, (int)(step * 1000.f + .5f), with_u);
    return PV_OK;
}
/* "flux sums" of a surface lookup = a function of the query point alone (so the image does not depend on batching or thread count):
 * Lr from the position, Lt = 0 */
int pv_surface_lphoton(pv_ctx *c, const float *pts, const float *nf, uint64_t n, uint32_t k, float r2, uint64_t paths, float *Lr, float *Lt) {
    (void)nf; (void)r2;
    for (uint64_t i = 0; i < n; ++i) {
        float v = 0.05f + 0.01f * (float)((int)(fabsf(pts[3 * i] * 7.f + pts[3 * i + 1] * 3.f + pts[3 * i + 2]) * 10.f) % 10);
        for (int b = 0; b < PV_NSPEC; ++b) { Lr[PV_NSPEC * i + b] = v * (c->selected == PV_MAP_CAUSTIC ? 1.f : 2.f); Lt[PV_NSPEC * i + b] = 0.f; }
    }
    logf_("surface_lphoton dev=%d map=%d n=%llu k=%u paths=%llu\n", c->device, c->selected, (unsigned long long)n, k, (unsigned long long)paths);
    return PV_OK;
}
/* "indirect radiance" of a final-gather ray = a hash of its global index, like fake_li */
int pv_final_gather(pv_ctx *c, const pv_ray *r, uint64_t n, float step, uint64_t seed, uint64_t base, float *L, uint32_t *idx) {
    (void)r; (void)step; (void)seed;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t g = base + i;
        uint32_t h = (uint32_t)(g * 2246822519u) ^ (uint32_t)(g >> 32);
        for (int b = 0; b < PV_NSPEC; ++b) L[PV_NSPEC * i + b] = (float)((h >> (b % 16)) & 0xffu) / 2550.f;
        if (idx) idx[i] = (uint32_t)(h % 300u);
    }
    logf_("final_gather dev=%d base=%llu n=%llu\n", c->device, (unsigned long long)base, (unsigned long long)n);
    return PV_OK;
}
