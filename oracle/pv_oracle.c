/* pv_oracle.c -- TEST INFRASTRUCTURE ONLY (see pv_oracle.h).
 *
 * Plain-C restatement of the reference's volumetric photon-mapping path.
 * Every function cites the reference file:line it follows.  Arithmetic is
 * unfused IEEE fp32 in the reference's operation order (compile with
 * -ffp-contract=off), so MT-mode results agree with the reference binary to
 * the last bit wherever only + - * / sqrt and glibc libm are involved.
 *
 * Parity status: PINNED against oracle/_ref/ref_harness output
 * (tests/golden/, tests/test_oracle_golden.py).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "pv_oracle.h"
#include "../include/pv_rng.h"

#define NS PV_NSPEC
#define PI_F 3.14159265358979323846f      /* core/pbrt.h:188 */
#define INV_PI_F 0.31830988618379067154f
#define ONE_MINUS_EPS 0.99999994f          /* core/montecarlo.h:50 */

typedef struct { float x, y, z; } v3;
typedef struct { float c[NS]; } spec;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, float f) { return V(f * a.x, f * a.y, f * a.z); }
/* core/geometry.h:92-96: Vector/f multiplies by the reciprocal */
static inline v3 vdiv(v3 a, float f) { float inv = 1.f / f; return V(a.x * inv, a.y * inv, a.z * inv); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float vlen(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline v3 vnorm(v3 a) { return vdiv(a, vlen(a)); }
/* core/geometry.h:477-484: products and differences in double, rounded once */
static inline v3 vcross(v3 a, v3 b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)), (float)((ax * by) - (ay * bx)));
}
static inline v3 ray_at(v3 o, v3 d, float t) { return vadd(o, vmul(d, t)); }
static inline float dist2(v3 a, v3 b) { v3 d = vsub(a, b); return d.x * d.x + d.y * d.y + d.z * d.z; }
static inline float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* core/transform.h:192-212 (points), :215-221 (vectors) */
static inline v3 xf_point(const float *m, v3 p) {
    float x = p.x, y = p.y, z = p.z;
    float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (wp == 1.f) return V(xp, yp, zp);
    float inv = 1.f / wp;
    return V(inv * xp, inv * yp, inv * zp);
}
static inline v3 xf_vec(const float *m, v3 v) {
    float x = v.x, y = v.y, z = v.z;
    return V(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}

/* ---------------------------------------------------------------- spectra */
static inline spec s_const(float v) { spec r; for (int i = 0; i < NS; ++i) r.c[i] = v; return r; }
static inline spec s_load(const float *p) { spec r; memcpy(r.c, p, sizeof(r.c)); return r; }
static inline int s_black(const spec *s) { for (int i = 0; i < NS; ++i) if (s->c[i] != 0.f) return 0; return 1; }
/* core/spectrum.h:433-439 */
static inline float s_y(const pv_scene_desc *sc, const spec *s) {
    float yy = 0.f;
    for (int i = 0; i < NS; ++i) yy += sc->cie_y[i] * s->c[i];
    return yy * (float)(700 - 400) / (float)(106.856895f * NS);
}
float pvo_spectrum_y(const pv_scene_desc *sc, const float *c) { spec s = s_load(c); return s_y(sc, &s); }

/* ---------------------------------------------------------------- MT19937 (core/rng.cpp:43-107) */
typedef struct { uint32_t mt[624]; int mti; } mt_rng;
static void mt_seed(mt_rng *r, uint32_t seed) {
    r->mt[0] = seed;
    for (r->mti = 1; r->mti < 624; r->mti++)
        r->mt[r->mti] = 1812433253u * (r->mt[r->mti - 1] ^ (r->mt[r->mti - 1] >> 30)) + (uint32_t)r->mti;
}
static uint32_t mt_next(mt_rng *r) {
    static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
    uint32_t y;
    if (r->mti >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
            r->mt[kk] = r->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < 623; kk++) {
            y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
            r->mt[kk] = r->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (r->mt[623] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
        r->mt[623] = r->mt[396] ^ (y >> 1) ^ mag01[y & 1u];
        r->mti = 0;
    }
    y = r->mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
uint32_t pvo_mt_first(uint32_t seed, uint32_t *out, uint32_t n) {
    mt_rng r; mt_seed(&r, seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = mt_next(&r);
    return n;
}
void pvo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    pv_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
uint32_t pvo_permute(uint32_t i, uint32_t l, uint32_t p) { return pv_permute(i, l, p); }
float pvo_van_der_corput(uint32_t n, uint32_t scramble) { return pv_van_der_corput(n, scramble); }

/* a sequential stream: either the shared MT or a per-path Philox stream */
typedef struct {
    int mode;
    mt_rng *mt;
    uint32_t k0, k1, c0, c1, j, pos, buf[4];
} stream;
static uint32_t st_u32(stream *s) {
    if (s->mode == PVO_RNG_MT) return mt_next(s->mt);
    if (s->pos == 4) { pv_philox4x32_10(s->c0, s->c1, s->j++, PV_RNG_PATH, s->k0, s->k1, s->buf); s->pos = 0; }
    return s->buf[s->pos++];
}
static float st_float(stream *s) { return pv_u32_to_float(st_u32(s)); }

/* ---------------------------------------------------------------- permuted Halton
 * core/montecarlo.h:206-243, core/montecarlo.cpp:380-397, Shuffle :175-181 */
typedef struct { uint32_t perm[2 + 3 + 5 + 7 + 11 + 13]; } halton6;
static const uint32_t halton_base[6] = {2, 3, 5, 7, 11, 13};
static void halton6_init(halton6 *h, mt_rng *rng) {
    uint32_t *p = h->perm;
    for (int d = 0; d < 6; ++d) {
        uint32_t b = halton_base[d];
        for (uint32_t i = 0; i < b; ++i) p[i] = i;
        for (uint32_t i = 0; i < b; ++i) {
            uint32_t other = i + (mt_next(rng) % (b - i));
            uint32_t t = p[i]; p[i] = p[other]; p[other] = t;
        }
        p += b;
    }
}
static void halton6_sample(const halton6 *h, uint32_t n0, float out[6]) {
    const uint32_t *p = h->perm;
    for (int d = 0; d < 6; ++d) {
        uint32_t base = halton_base[d], n = n0;
        double val = 0, invBase = 1. / base, invBi = invBase;
        while (n > 0) {
            uint32_t d_i = p[n % base];
            val += d_i * invBi;
            n = (uint32_t)(n * invBase);     /* `n *= invBase` on a uint32_t */
            invBi *= invBase;
        }
        float f = (float)val;
        out[d] = f < ONE_MINUS_EPS ? f : ONE_MINUS_EPS;
        p += base;
    }
}
void pvo_halton6(uint32_t seed, uint32_t n, float out[6]) {
    mt_rng r; mt_seed(&r, seed);
    halton6 h; halton6_init(&h, &r);
    halton6_sample(&h, n, out);
}

/* ---------------------------------------------------------------- sampling maps (core/montecarlo.cpp) */
static v3 uniform_sample_sphere(float u1, float u2) {           /* :283-290 */
    float z = 1.f - 2.f * u1;
    float r = sqrtf(fmaxf(0.f, 1.f - z * z));
    float phi = 2.f * PI_F * u2;
    return V(r * cosf(phi), r * sinf(phi), z);
}
static v3 uniform_sample_cone(float u1, float u2, float costhetamax) {   /* :405-410 */
    float costheta = (1.f - u1) + u1 * costhetamax;
    float sintheta = sqrtf(1.f - costheta * costheta);
    float phi = u2 * 2.f * PI_F;
    return V(cosf(phi) * sintheta, sinf(phi) * sintheta, costheta);
}
static void concentric_sample_disk(float u1, float u2, float *dx, float *dy) {  /* :306-348 */
    float r, theta;
    float sx = 2 * u1 - 1, sy = 2 * u2 - 1;
    if (sx == 0.0 && sy == 0.0) { *dx = 0.0; *dy = 0.0; return; }
    if (sx >= -sy) {
        if (sx > sy) { r = sx; if (sy > 0.0) theta = sy / r; else theta = 8.0f + sy / r; }
        else { r = sy; theta = 2.0f - sx / r; }
    } else {
        if (sx <= sy) { r = -sx; theta = 4.0f - sy / r; }
        else { r = -sy; theta = 6.0f + sx / r; }
    }
    theta *= PI_F / 4.f;
    *dx = r * cosf(theta);
    *dy = r * sinf(theta);
}
static void coordinate_system(v3 v1, v3 *v2, v3 *v3o) {           /* core/geometry.h:508-518 */
    if (fabsf(v1.x) > fabsf(v1.y)) {
        float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
        *v2 = V(-v1.z * invLen, 0.f, v1.x * invLen);
    } else {
        float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
        *v2 = V(0.f, v1.z * invLen, -v1.y * invLen);
    }
    *v3o = vcross(v1, *v2);
}

/* ---------------------------------------------------------------- BBox (core/geometry.cpp:68-86, geometry.h:404-439) */
static int bbox_intersectp(const float *p0, const float *p1, v3 o, v3 d, float mint, float maxt, float *ht0, float *ht1) {
    float t0 = mint, t1 = maxt;
    for (int i = 0; i < 3; ++i) {
        float invRayDir = 1.f / comp(d, i);
        float tNear = (p0[i] - comp(o, i)) * invRayDir;
        float tFar = (p1[i] - comp(o, i)) * invRayDir;
        if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; }
        t0 = tNear > t0 ? tNear : t0;
        t1 = tFar < t1 ? tFar : t1;
        if (t0 > t1) return 0;
    }
    *ht0 = t0; *ht1 = t1;
    return 1;
}
static int bbox_inside(const float *p0, const float *p1, v3 p) {
    return p.x >= p0[0] && p.x <= p1[0] && p.y >= p0[1] && p.y <= p1[1] && p.z >= p0[2] && p.z <= p1[2];
}

/* ---------------------------------------------------------------- media */
typedef struct { uint64_t density_samples; } med_counters;

static inline float lerpf(float t, float a, float b) { return (1.f - t) * a + t * b; }   /* core/pbrt.h:218 */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline float grid_D(const pv_medium *m, int x, int y, int z) {                      /* volumes/volumegrid.h:60-65 */
    x = clampi(x, 0, m->nx - 1); y = clampi(y, 0, m->ny - 1); z = clampi(z, 0, m->nz - 1);
    return m->density[(size_t)z * m->nx * m->ny + (size_t)y * m->nx + x];
}
/* volumes/volumegrid.cpp:39-57 */
static float grid_density(const pv_medium *m, v3 Pobj, med_counters *mc) {
    if (mc) mc->density_samples++;
    if (!bbox_inside(m->p0, m->p1, Pobj)) return 0;
    if (m->type == PV_MEDIUM_EXPONENTIAL) {           /* volumes/exponential.h:57-61; density = {a, b, updir.xyz} */
        float height = vdot(vsub(Pobj, V(m->p0[0], m->p0[1], m->p0[2])), V(m->density[2], m->density[3], m->density[4]));
        return m->density[0] * expf(-m->density[1] * height);
    }
    v3 vox = V((Pobj.x - m->p0[0]) / (m->p1[0] - m->p0[0]), (Pobj.y - m->p0[1]) / (m->p1[1] - m->p0[1]),
               (Pobj.z - m->p0[2]) / (m->p1[2] - m->p0[2]));
    vox.x = vox.x * m->nx - .5f; vox.y = vox.y * m->ny - .5f; vox.z = vox.z * m->nz - .5f;
    int vx = (int)floorf(vox.x), vy = (int)floorf(vox.y), vz = (int)floorf(vox.z);
    float dx = vox.x - vx, dy = vox.y - vy, dz = vox.z - vz;
    float d00 = lerpf(dx, grid_D(m, vx, vy, vz), grid_D(m, vx + 1, vy, vz));
    float d10 = lerpf(dx, grid_D(m, vx, vy + 1, vz), grid_D(m, vx + 1, vy + 1, vz));
    float d01 = lerpf(dx, grid_D(m, vx, vy, vz + 1), grid_D(m, vx + 1, vy, vz + 1));
    float d11 = lerpf(dx, grid_D(m, vx, vy + 1, vz + 1), grid_D(m, vx + 1, vy + 1, vz + 1));
    float d0 = lerpf(dy, d00, d10);
    float d1 = lerpf(dy, d01, d11);
    return lerpf(dz, d0, d1);
}
static int med_is_homog(const pv_medium *m) { return m->type == PV_MEDIUM_HOMOGENEOUS || m->type == PV_MEDIUM_RAINBOW; }
/* VolumeRegion::IntersectP: volumes/homogeneous.h:58-61, volumes/volumegrid.h:56-59 */
static int med_intersectp(const pv_medium *m, v3 o, v3 d, float mint, float maxt, float *t0, float *t1) {
    v3 oo = xf_point(m->world_to_volume, o), dd = xf_vec(m->world_to_volume, d);
    return bbox_intersectp(m->p0, m->p1, oo, dd, mint, maxt, t0, t1);
}
/* sigma_a / sigma_s: volumes/homogeneous.h:62-70, core/volume.h:80-88 */
static spec med_sigma(const pv_medium *m, const float *sig, v3 p, med_counters *mc) {
    v3 po = xf_point(m->world_to_volume, p);
    spec r;
    if (med_is_homog(m)) {
        if (bbox_inside(m->p0, m->p1, po)) memcpy(r.c, sig, sizeof(r.c)); else r = s_const(0.f);
    } else {
        float dns = grid_density(m, po, mc);
        for (int i = 0; i < NS; ++i) r.c[i] = sig[i] * dns;
    }
    return r;
}
/* PhaseHG core/volume.cpp:150-154 */
static float phase_hg(v3 w, v3 wp, float g) {
    float costheta = vdot(w, wp);
    return 1.f / (4.f * PI_F) * (1.f - g * g) / powf(1.f + g * g - 2.f * g * costheta, 1.5f);
}
/* VolumeRegion::p: volumes/homogeneous.h:74-77 (gated by the extent), core/volume.h:92-94 (not gated) */
static float med_p(const pv_medium *m, v3 p, v3 w, v3 wp) {
    if (med_is_homog(m) && !bbox_inside(m->p0, m->p1, xf_point(m->world_to_volume, p))) return 0.f;
    return phase_hg(w, wp, m->g);
}
/* tau: volumes/homogeneous.h:78-82; core/volume.cpp:296-310 */
static spec med_tau(const pv_medium *m, v3 o, v3 d, float mint, float maxt, float stepSize, float u, med_counters *mc) {
    float t0, t1;
    spec tau = s_const(0.f);
    if (med_is_homog(m)) {
        if (!med_intersectp(m, o, d, mint, maxt, &t0, &t1)) return tau;
        v3 a = ray_at(o, d, t0), b = ray_at(o, d, t1);
        float dist = vlen(vsub(a, b));
        for (int i = 0; i < NS; ++i) tau.c[i] = (m->sigma_a[i] + m->sigma_s[i]) * dist;
        return tau;
    }
    float length = vlen(d);
    if (length == 0.f) return tau;
    v3 dn = vdiv(d, length);
    float nmint = mint * length, nmaxt = maxt * length;
    if (!med_intersectp(m, o, dn, nmint, nmaxt, &t0, &t1)) return tau;
    t0 += u * stepSize;
    while (t0 < t1) {
        v3 po = xf_point(m->world_to_volume, ray_at(o, dn, t0));
        float dns = grid_density(m, po, mc);
        for (int i = 0; i < NS; ++i) tau.c[i] += (m->sigma_a[i] + m->sigma_s[i]) * dns;
        t0 += stepSize;
    }
    for (int i = 0; i < NS; ++i) tau.c[i] = tau.c[i] * stepSize;
    return tau;
}
static spec s_exp_neg(const spec *tau) { spec r; for (int i = 0; i < NS; ++i) r.c[i] = expf(-tau->c[i]); return r; }

/* ---------------------------------------------------------------- AggregateVolume (core/volume.cpp:178-261)
 * What core/api.cpp:1200-1205 builds when a scene file has more than one Volume.  pv_scene_desc carries ONE medium (the device
 * path has no aggregate yet), so the further regions are handed to the oracle on the side: region 0 is
 * sc->medium, regions 1.. are pvo_set_more_media()'s.  With no further regions every vol_* below IS the single-medium
 * function (a scene with one Volume is never wrapped in an AggregateVolume). */
static const pv_medium *g_more_media = NULL;
static uint32_t g_n_more = 0;
void pvo_set_more_media(const pv_medium *more, uint32_t n) { g_more_media = n ? more : NULL; g_n_more = more ? n : 0; }
enum { SEL_A = 0, SEL_S = 1, SEL_LE = 2 };
static inline const float *med_sel(const pv_medium *m, int sel) { return sel == SEL_A ? m->sigma_a : (sel == SEL_S ? m->sigma_s : m->le); }
static inline const pv_medium *vol_region(const pv_scene_desc *sc, uint32_t k) { return k == 0 ? sc->medium : &g_more_media[k - 1]; }
static int vol_intersectp(const pv_scene_desc *sc, v3 o, v3 d, float mint, float maxt, float *t0, float *t1) {   /* :238-250 */
    if (!g_n_more) return med_intersectp(sc->medium, o, d, mint, maxt, t0, t1);
    *t0 = INFINITY; *t1 = -INFINITY;
    for (uint32_t k = 0; k <= g_n_more; ++k) {
        float tr0, tr1;
        if (med_intersectp(vol_region(sc, k), o, d, mint, maxt, &tr0, &tr1)) { *t0 = *t0 < tr0 ? *t0 : tr0; *t1 = *t1 > tr1 ? *t1 : tr1; }
    }
    return *t0 < *t1;
}
static spec vol_sigma(const pv_scene_desc *sc, int sel, v3 p, med_counters *mc) {                                  /* :185-207 */
    if (!g_n_more) return med_sigma(sc->medium, med_sel(sc->medium, sel), p, mc);
    spec s = s_const(0.f);
    for (uint32_t k = 0; k <= g_n_more; ++k) {
        const pv_medium *m = vol_region(sc, k);
        spec r = med_sigma(m, med_sel(m, sel), p, mc);
        for (int i = 0; i < NS; ++i) s.c[i] += r.c[i];
    }
    return s;
}
static float vol_p(const pv_scene_desc *sc, v3 p, v3 w, v3 wp) {                                                  /* :210-219 */
    if (!g_n_more) return med_p(sc->medium, p, w, wp);
    float ph = 0, sumWt = 0;
    for (uint32_t k = 0; k <= g_n_more; ++k) {
        const pv_medium *m = vol_region(sc, k);
        spec ss = med_sigma(m, m->sigma_s, p, NULL);
        float wt = s_y(sc, &ss);
        sumWt += wt;
        ph += wt * med_p(m, p, w, wp);
    }
    return ph / sumWt;
}
static spec vol_tau(const pv_scene_desc *sc, v3 o, v3 d, float mint, float maxt, float stepSize, float u, med_counters *mc) {   /* :230-235 */
    if (!g_n_more) return med_tau(sc->medium, o, d, mint, maxt, stepSize, u, mc);
    spec t = s_const(0.f);
    for (uint32_t k = 0; k <= g_n_more; ++k) {
        spec r = med_tau(vol_region(sc, k), o, d, mint, maxt, stepSize, u, mc);
        for (int i = 0; i < NS; ++i) t.c[i] += r.c[i];
    }
    return t;
}

int pvo_transmittance(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n, float step, const float *offset_u, float *T) {
    for (uint64_t i = 0; i < n; ++i) {
        spec tr = s_const(1.f);
        if (sc->medium && sc->medium->type != PV_MEDIUM_NONE) {
            const pv_ray *r = &rays[i];
            spec tau = vol_tau(sc, V(r->o[0], r->o[1], r->o[2]), V(r->d[0], r->d[1], r->d[2]), r->mint, r->maxt,
                               step, offset_u ? offset_u[i] : 0.5f, NULL);
            tr = s_exp_neg(&tau);
        }
        memcpy(T + NS * i, tr.c, sizeof(tr.c));
    }
    return 0;
}

/* ---------------------------------------------------------------- BVH + triangles */
typedef struct { uint64_t nodes_visited, tri_tests; } bvh_counters;

/* accelerators/bvh.cpp:167-189 */
static inline int node_slab(const pv_bvh_node *nd, v3 o, float mint, float maxt, v3 invDir, const uint32_t dirIsNeg[3]) {
    const float *b = nd->bounds;
    float tmin = (b[dirIsNeg[0] ? 3 : 0] - o.x) * invDir.x;
    float tmax = (b[dirIsNeg[0] ? 0 : 3] - o.x) * invDir.x;
    float tymin = (b[dirIsNeg[1] ? 4 : 1] - o.y) * invDir.y;
    float tymax = (b[dirIsNeg[1] ? 1 : 4] - o.y) * invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return 0;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (b[dirIsNeg[2] ? 5 : 2] - o.z) * invDir.z;
    float tzmax = (b[dirIsNeg[2] ? 2 : 5] - o.z) * invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return 0;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < maxt) && (tmax > mint);
}
/* shapes/trianglemesh.cpp:127-158 (Intersect) == :211-241 (IntersectP) up to the t test */
static inline int tri_hit(const float *tv, v3 o, v3 d, float mint, float maxt, float *tHit) {
    v3 p1 = V(tv[0], tv[1], tv[2]), p2 = V(tv[3], tv[4], tv[5]), p3 = V(tv[6], tv[7], tv[8]);
    v3 e1 = vsub(p2, p1), e2 = vsub(p3, p1);
    v3 s1 = vcross(d, e2);
    float divisor = vdot(s1, e1);
    if (divisor == 0.f) return 0;
    float invDivisor = 1.f / divisor;
    v3 s = vsub(o, p1);
    float b1 = vdot(s, s1) * invDivisor;
    if (b1 < 0.f || b1 > 1.f) return 0;
    v3 s2 = vcross(s, e1);
    float b2 = vdot(d, s2) * invDivisor;
    if (b2 < 0.f || b1 + b2 > 1.f) return 0;
    float t = vdot(e2, s2) * invDivisor;
    if (t < mint || t > maxt) return 0;
    *tHit = t;
    return 1;
}
/* shapes/sphere.cpp:58-110 (Intersect) == :167-214 (IntersectP): the hit parameter only.  Quadratic: core/pbrt.h:309-323 */
static inline void sphere_phit(const pv_sphere *s, v3 ro, v3 rd, float thit, v3 *phit, float *phi) {
    *phit = ray_at(ro, rd, thit);
    if (phit->x == 0.f && phit->y == 0.f) phit->x = 1e-5f * s->radius;
    float ph = atan2f(phit->y, phit->x);
    if (ph < 0.) ph = (float)((double)ph + (double)2.f * 3.14159265358979323846);
    *phi = ph;
}
static inline int sphere_clipped(const pv_sphere *s, v3 phit, float phi) {
    return (s->zmin > -s->radius && phit.z < s->zmin) || (s->zmax < s->radius && phit.z > s->zmax) || phi > s->phi_max;
}
static int sphere_hit(const pv_sphere *s, v3 o, v3 d, float mint, float maxt, float *tHit) {
    v3 ro = xf_point(s->world_to_object, o), rd = xf_vec(s->world_to_object, d);
    float A = rd.x * rd.x + rd.y * rd.y + rd.z * rd.z;
    float B = 2 * (rd.x * ro.x + rd.y * ro.y + rd.z * ro.z);
    float C = ro.x * ro.x + ro.y * ro.y + ro.z * ro.z - s->radius * s->radius;
    float discrim = B * B - 4.f * A * C;
    if (discrim < 0.) return 0;
    float rootDiscrim = sqrtf(discrim);
    float q = B < 0 ? -.5f * (B - rootDiscrim) : -.5f * (B + rootDiscrim);
    float t0 = q / A, t1 = C / q;
    if (t0 > t1) { float t = t0; t0 = t1; t1 = t; }
    if (t0 > maxt || t1 < mint) return 0;
    float thit = t0;
    if (t0 < mint) { thit = t1; if (thit > maxt) return 0; }
    v3 phit; float phi;
    sphere_phit(s, ro, rd, thit, &phit, &phi);
    if (sphere_clipped(s, phit, phi)) {
        if (thit == t1) return 0;
        if (t1 > maxt) return 0;
        thit = t1;
        sphere_phit(s, ro, rd, thit, &phit, &phi);
        if (sphere_clipped(s, phit, phi)) return 0;
    }
    *tHit = thit;
    return 1;
}
static inline int prim_hit(const pv_scene_desc *sc, uint32_t prim, v3 o, v3 d, float mint, float maxt, float *tHit) {
    if (sc->n_spheres && sc->prim_shape[prim] != PV_SHAPE_TRIANGLE) return sphere_hit(&sc->spheres[sc->prim_shape[prim]], o, d, mint, maxt, tHit);
    return tri_hit(sc->tri_verts + 9 * (size_t)prim, o, d, mint, maxt, tHit);
}
/* accelerators/bvh.cpp:585-636; returns primitive index (reordered array) or -1; *maxt shrinks */
static int bvh_intersect(const pv_scene_desc *sc, v3 o, v3 d, float mint, float *maxt, bvh_counters *bc) {
    if (!sc->n_nodes) return -1;
    int hit = -1;
    v3 invDir = V(1.f / d.x, 1.f / d.y, 1.f / d.z);
    uint32_t dirIsNeg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
    uint32_t todoOffset = 0, nodeNum = 0, todo[64];
    for (;;) {
        const pv_bvh_node *node = &sc->nodes[nodeNum];
        if (bc) bc->nodes_visited++;
        if (node_slab(node, o, mint, *maxt, invDir, dirIsNeg)) {
            if (node->n_primitives > 0) {
                for (uint32_t i = 0; i < node->n_primitives; ++i) {
                    float t;
                    if (bc) bc->tri_tests++;
                    if (prim_hit(sc, node->offset + i, o, d, mint, *maxt, &t)) {
                        hit = (int)(node->offset + i);
                        *maxt = t;
                    }
                }
                if (todoOffset == 0) break;
                nodeNum = todo[--todoOffset];
            } else {
                if (dirIsNeg[node->axis]) { todo[todoOffset++] = nodeNum + 1; nodeNum = node->offset; }
                else { todo[todoOffset++] = node->offset; nodeNum = nodeNum + 1; }
            }
        } else {
            if (todoOffset == 0) break;
            nodeNum = todo[--todoOffset];
        }
    }
    return hit;
}
/* accelerators/bvh.cpp:639-685 */
static int bvh_intersectp(const pv_scene_desc *sc, v3 o, v3 d, float mint, float maxt, bvh_counters *bc) {
    if (!sc->n_nodes) return 0;
    v3 invDir = V(1.f / d.x, 1.f / d.y, 1.f / d.z);
    uint32_t dirIsNeg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
    uint32_t todoOffset = 0, nodeNum = 0, todo[64];
    for (;;) {
        const pv_bvh_node *node = &sc->nodes[nodeNum];
        if (bc) bc->nodes_visited++;
        if (node_slab(node, o, mint, maxt, invDir, dirIsNeg)) {
            if (node->n_primitives > 0) {
                for (uint32_t i = 0; i < node->n_primitives; ++i) {
                    float t;
                    if (bc) bc->tri_tests++;
                    if (prim_hit(sc, node->offset + i, o, d, mint, maxt, &t)) return 1;
                }
                if (todoOffset == 0) break;
                nodeNum = todo[--todoOffset];
            } else {
                if (dirIsNeg[node->axis]) { todo[todoOffset++] = nodeNum + 1; nodeNum = node->offset; }
                else { todo[todoOffset++] = node->offset; nodeNum = nodeNum + 1; }
            }
        } else {
            if (todoOffset == 0) break;
            nodeNum = todo[--todoOffset];
        }
    }
    return 0;
}
int pvo_intersect(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n, uint32_t *prim, float *t) {
    for (uint64_t i = 0; i < n; ++i) {
        const pv_ray *r = &rays[i];
        float maxt = r->maxt;
        int h = bvh_intersect(sc, V(r->o[0], r->o[1], r->o[2]), V(r->d[0], r->d[1], r->d[2]), r->mint, &maxt, NULL);
        prim[i] = h < 0 ? 0xFFFFFFFFu : (uint32_t)h;
        t[i] = h < 0 ? INFINITY : maxt;
    }
    return 0;
}
int pvo_occluded(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n, uint8_t *hit) {
    for (uint64_t i = 0; i < n; ++i) {
        const pv_ray *r = &rays[i];
        hit[i] = (uint8_t)bvh_intersectp(sc, V(r->o[0], r->o[1], r->o[2]), V(r->d[0], r->d[1], r->d[2]), r->mint, r->maxt, NULL);
    }
    return 0;
}

/* ---------------------------------------------------------------- kd-tree (core/kdtree.h) */
typedef struct { float splitPos; uint32_t splitAxis, hasLeftChild, rightChild; } kdnode;
struct pvo_kdtree {
    uint32_t nNodes, nextFreeNode;
    kdnode *nodes;
    v3 *nodePos;          /* nodeData[].p  */
    uint32_t *nodeOrig;   /* node -> original photon index */
    const float *pos;     /* build input (not owned) */
};
#define KD_NOCHILD ((1u << 29) - 1)

/* CompareNode kdtree.h:87-94: (p[axis], pointer) -- pointer order == input index order */
static inline int kd_less(const float *pos, int axis, uint32_t a, uint32_t b) {
    float pa = pos[3 * (size_t)a + axis], pb = pos[3 * (size_t)b + axis];
    return pa == pb ? (a < b) : (pa < pb);
}
/* std::nth_element only guarantees the element at nth and the partition; the
 * comparator is a strict total order, so the resulting tree is unique. */
static void kd_nth(const float *pos, int axis, uint32_t *a, int lo, int hi, int nth) {
    while (hi - lo > 1) {
        uint32_t piv = a[lo + (hi - lo) / 2];
        int i = lo, j = hi - 1;
        while (i <= j) {
            while (kd_less(pos, axis, a[i], piv)) ++i;
            while (kd_less(pos, axis, piv, a[j])) --j;
            if (i <= j) { uint32_t t = a[i]; a[i] = a[j]; a[j] = t; ++i; --j; }
        }
        if (nth <= j) hi = j + 1;
        else if (nth >= i) lo = i;
        else return;
    }
}
static void kd_build(pvo_kdtree *t, uint32_t nodeNum, int start, int end, uint32_t *bn) {   /* kdtree.h:113-147 */
    const float *pos = t->pos;
    if (start + 1 == end) {
        t->nodes[nodeNum].splitAxis = 3; t->nodes[nodeNum].rightChild = KD_NOCHILD; t->nodes[nodeNum].hasLeftChild = 0;
        t->nodePos[nodeNum] = V(pos[3 * (size_t)bn[start]], pos[3 * (size_t)bn[start] + 1], pos[3 * (size_t)bn[start] + 2]);
        t->nodeOrig[nodeNum] = bn[start];
        return;
    }
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = start; i < end; ++i)
        for (int a = 0; a < 3; ++a) {
            float v = pos[3 * (size_t)bn[i] + a];
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    int splitAxis = (dx > dy && dx > dz) ? 0 : (dy > dz ? 1 : 2);    /* BBox::MaximumExtent geometry.h:424-432 */
    int splitPos = (start + end) / 2;
    kd_nth(pos, splitAxis, bn, start, end, splitPos);
    kdnode *nd = &t->nodes[nodeNum];
    nd->splitPos = pos[3 * (size_t)bn[splitPos] + splitAxis];
    nd->splitAxis = (uint32_t)splitAxis; nd->rightChild = KD_NOCHILD; nd->hasLeftChild = 0;
    t->nodePos[nodeNum] = V(pos[3 * (size_t)bn[splitPos]], pos[3 * (size_t)bn[splitPos] + 1], pos[3 * (size_t)bn[splitPos] + 2]);
    t->nodeOrig[nodeNum] = bn[splitPos];
    if (start < splitPos) {
        t->nodes[nodeNum].hasLeftChild = 1;
        uint32_t childNum = t->nextFreeNode++;
        kd_build(t, childNum, start, splitPos, bn);
    }
    if (splitPos + 1 < end) {
        t->nodes[nodeNum].rightChild = t->nextFreeNode++;
        kd_build(t, t->nodes[nodeNum].rightChild, splitPos + 1, end, bn);
    }
}
pvo_kdtree *pvo_kdtree_build(const float *pos, uint64_t n) {
    pvo_kdtree *t = (pvo_kdtree *)calloc(1, sizeof(*t));
    t->nNodes = (uint32_t)n; t->nextFreeNode = 1; t->pos = pos;
    if (!n) return t;
    t->nodes = (kdnode *)malloc(sizeof(kdnode) * n);
    t->nodePos = (v3 *)malloc(sizeof(v3) * n);
    t->nodeOrig = (uint32_t *)malloc(sizeof(uint32_t) * n);
    uint32_t *bn = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint64_t i = 0; i < n; ++i) bn[i] = (uint32_t)i;
    kd_build(t, 0, 0, (int)n, bn);
    free(bn);
    t->pos = NULL;
    return t;
}
void pvo_kdtree_free(pvo_kdtree *t) {
    if (!t) return;
    free(t->nodes); free(t->nodePos); free(t->nodeOrig); free(t);
}

/* ClosePhoton + PhotonProcess (core/photonshooter.h:40-50,186-203); the photon
 * "pointer" is the kd-tree node index (nodeData order). */
typedef struct { uint32_t node; float d2; } closeph;
static inline int cp_less(closeph a, closeph b) { return a.d2 == b.d2 ? (a.node < b.node) : (a.d2 < b.d2); }
static void heap_sift_down(closeph *h, uint32_t n, uint32_t i) {
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cp_less(h[m], h[l])) m = l;
        if (r < n && cp_less(h[m], h[r])) m = r;
        if (m == i) return;
        closeph t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}
typedef struct { closeph *photons; uint32_t nLookup, nFound; uint64_t rejected_at_bound; } photon_proc;
static void proc_call(photon_proc *pr, uint32_t node, float d2, float *maxDistSquared) {
    if (pr->nFound < pr->nLookup) {
        pr->photons[pr->nFound].node = node; pr->photons[pr->nFound].d2 = d2; pr->nFound++;
        if (pr->nFound == pr->nLookup) {
            for (int i = (int)pr->nLookup / 2 - 1; i >= 0; --i) heap_sift_down(pr->photons, pr->nLookup, (uint32_t)i);
            *maxDistSquared = pr->photons[0].d2;
        }
    } else {
        /* pop_heap + overwrite last + push_heap == replace the max and restore */
        pr->photons[0].node = node; pr->photons[0].d2 = d2;
        heap_sift_down(pr->photons, pr->nLookup, 0);
        *maxDistSquared = pr->photons[0].d2;
    }
}
static void kd_lookup(const pvo_kdtree *t, uint32_t nodeNum, v3 p, photon_proc *pr, float *maxDistSquared) {   /* kdtree.h:157-183 */
    const kdnode *node = &t->nodes[nodeNum];
    int axis = (int)node->splitAxis;
    if (axis != 3) {
        float pa = comp(p, axis);
        float d2 = (pa - node->splitPos) * (pa - node->splitPos);
        if (pa <= node->splitPos) {
            if (node->hasLeftChild) kd_lookup(t, nodeNum + 1, p, pr, maxDistSquared);
            if (d2 < *maxDistSquared && node->rightChild < t->nNodes) kd_lookup(t, node->rightChild, p, pr, maxDistSquared);
        } else {
            if (node->rightChild < t->nNodes) kd_lookup(t, node->rightChild, p, pr, maxDistSquared);
            if (d2 < *maxDistSquared && node->hasLeftChild) kd_lookup(t, nodeNum + 1, p, pr, maxDistSquared);
        }
    }
    float d2 = dist2(t->nodePos[nodeNum], p);
    if (d2 < *maxDistSquared) proc_call(pr, nodeNum, d2, maxDistSquared);
    else if (d2 == *maxDistSquared && pr->nFound == pr->nLookup) pr->rejected_at_bound++;
}
typedef struct { uint32_t idx; float d2; } idxd2;
static int idxd2_cmp(const void *a, const void *b) {
    const idxd2 *x = (const idxd2 *)a, *y = (const idxd2 *)b;
    if (x->d2 != y->d2) return x->d2 < y->d2 ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
int pvo_knn(const pvo_kdtree *t, const float *pts, uint64_t n, uint32_t k, float r2, uint32_t *idx, float *d2,
            uint32_t *nfound, uint64_t *boundary_ties) {
    closeph *buf = (closeph *)malloc(sizeof(closeph) * (k ? k : 1));
    idxd2 *res = (idxd2 *)malloc(sizeof(idxd2) * (k ? k : 1));
    uint64_t ties = 0;
    for (uint64_t q = 0; q < n; ++q) {
        photon_proc pr = {buf, k, 0, 0};
        float md2 = r2;
        if (t->nNodes && k) kd_lookup(t, 0, V(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]), &pr, &md2);
        if (pr.rejected_at_bound) ties++;
        for (uint32_t j = 0; j < pr.nFound; ++j) { res[j].idx = t->nodeOrig[buf[j].node]; res[j].d2 = buf[j].d2; }
        qsort(res, pr.nFound, sizeof(idxd2), idxd2_cmp);
        for (uint32_t j = 0; j < k; ++j) {
            idx[q * k + j] = j < pr.nFound ? res[j].idx : 0xFFFFFFFFu;
            d2[q * k + j] = j < pr.nFound ? res[j].d2 : INFINITY;
        }
        nfound[q] = pr.nFound;
    }
    if (boundary_ties) *boundary_ties = ties;
    free(buf); free(res);
    return 0;
}
int pvo_knn_brute(const float *pos, uint64_t nph, const float *pts, uint64_t n, uint32_t k, float r2,
                  uint32_t *idx, float *d2, uint32_t *nfound) {
    idxd2 *cand = (idxd2 *)malloc(sizeof(idxd2) * (nph ? nph : 1));
    for (uint64_t q = 0; q < n; ++q) {
        v3 p = V(pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]);
        uint64_t m = 0;
        for (uint64_t i = 0; i < nph; ++i) {
            float dd = dist2(V(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), p);
            if (dd < r2) { cand[m].idx = (uint32_t)i; cand[m].d2 = dd; ++m; }
        }
        qsort(cand, m, sizeof(idxd2), idxd2_cmp);
        uint32_t nf = (uint32_t)(m < k ? m : k);
        for (uint32_t j = 0; j < k; ++j) {
            idx[q * k + j] = j < nf ? cand[j].idx : 0xFFFFFFFFu;
            d2[q * k + j] = j < nf ? cand[j].d2 : INFINITY;
        }
        nfound[q] = nf;
    }
    free(cand);
    return 0;
}

/* ---------------------------------------------------------------- LPhoton (integrators/photonvolume.cpp:65-108) */
static spec lphoton(const pv_scene_desc *sc, const pvo_kdtree *t, const float *wi, const float *alpha, v3 w, v3 pt,
                    uint32_t nLookup, float maxDistSquare, closeph *buf, pv_gather_stats *st) {
    spec L = s_const(0.f);
    if (!t || !t->nNodes) return L;
    photon_proc pr = {buf, nLookup, 0, 0};
    kd_lookup(t, 0, pt, &pr, &maxDistSquare);
    int nFound = (int)pr.nFound;
    if (st) { st->lookups++; st->photons_found += pr.nFound; if (pr.nFound == nLookup) st->heap_lookups++; }
    if (nFound < 10) return L;
    spec totalFlux = s_const(0.f);
    float maxmd = 0.0f;
    v3 nw = vneg(w);
    for (int i = 0; i < nFound; ++i) {
        uint32_t node = buf[i].node, orig = t->nodeOrig[node];
        float distSq = buf[i].d2;
        if (distSq > maxmd) maxmd = distSq;
        v3 pwi = V(wi[3 * (size_t)orig], wi[3 * (size_t)orig + 1], wi[3 * (size_t)orig + 2]);
        float ph = vol_p(sc, t->nodePos[node], pwi, nw);
        const float *a = alpha + NS * (size_t)orig;
        for (int b = 0; b < NS; ++b) totalFlux.c[b] += a[b] * ph;
    }
    float distSq = maxmd;
    float dV = distSq * sqrtf(distSq);
    spec scale = vol_sigma(sc, SEL_S, pt, NULL);
    if (dV != 0.0 && !s_black(&scale)) {
        float f = (float)(4.0 / 3.0 * (double)PI_F * (double)dV);      /* double expr, converted at `float * Spectrum` */
        for (int b = 0; b < NS; ++b) L.c[b] += totalFlux.c[b] / (scale.c[b] * f);
    }
    return L;
}
int pvo_lphoton(const pv_scene_desc *sc, const pvo_kdtree *t, const float *wi, const float *alpha, const float *pts,
                const float *w, uint64_t n, uint32_t nused, float maxdist, float *L) {
    closeph *buf = (closeph *)malloc(sizeof(closeph) * (nused ? nused : 1));
    for (uint64_t i = 0; i < n; ++i) {
        spec l = lphoton(sc, t, wi, alpha, V(w[3 * i], w[3 * i + 1], w[3 * i + 2]), V(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]),
                         nused, maxdist * maxdist, buf, NULL);
        memcpy(L + NS * i, l.c, sizeof(l.c));
    }
    free(buf);
    return 0;
}

/* ---------------------------------------------------------------- lights, point-query Sample_L */
typedef struct { v3 o, d; float mint, maxt; } visray;
static float spot_falloff(const pv_light *l, v3 w) {                 /* lights/spot.cpp:60-69 */
    v3 wl = vnorm(xf_vec(l->world_to_light, w));
    float costheta = wl.z;
    if (costheta < l->cos_total_width) return 0.f;
    if (costheta > l->cos_falloff_start) return 1.f;
    float delta = (costheta - l->cos_total_width) / (l->cos_falloff_start - l->cos_total_width);
    return delta * delta * delta * delta;
}
/* lights/point.cpp:50-57, lights/spot.cpp:50-57, lights/distant.cpp:48-55; VisibilityTester core/light.h:85-101 */
static spec light_sample_L_point(const pv_light *l, v3 p, v3 *wi, float *pdf, visray *vis) {
    spec L;
    *pdf = 1.f;
    if (l->type == PV_LIGHT_DISTANT) {
        *wi = V(l->dir[0], l->dir[1], l->dir[2]);
        vis->o = p; vis->d = *wi; vis->mint = 0.f; vis->maxt = INFINITY;
        return s_load(l->intensity);
    }
    v3 lp = V(l->pos[0], l->pos[1], l->pos[2]);
    *wi = vnorm(vsub(lp, p));
    float dist = vlen(vsub(p, lp));
    vis->o = p; vis->d = vdiv(vsub(lp, p), dist); vis->mint = 0.f; vis->maxt = dist * (1.f - 0.f);
    float ds = dist2(lp, p);
    if (l->type == PV_LIGHT_SPOT) {
        float fo = spot_falloff(l, vneg(*wi));
        for (int i = 0; i < NS; ++i) L.c[i] = (l->intensity[i] * fo) / ds;
    } else {
        for (int i = 0; i < NS; ++i) L.c[i] = l->intensity[i] / ds;
    }
    return L;
}

/* ---------------------------------------------------------------- DiffuseAreaLight over triangles (lights/diffuse.cpp:69-86,
 * ShapeSet core/light.cpp:114-172, Shape::Pdf core/shape.cpp:86-99, Triangle::Sample shapes/trianglemesh.cpp:444-456)
 * pv_light has no area-light fields yet: such a light keeps its slot in sc->lights[] as a placeholder of type
 * PVO_LIGHT_SLOT and its data comes in on the side (pvo_set_area_lights).  Point-query sampling only (the direct term of the
 * volume integrators); MT stream only (the Philox assignment of the three light-sample numbers is a device-side decision). */
static const pvo_area_light *g_area = NULL;
static uint32_t g_n_area = 0;
void pvo_set_area_lights(const pvo_area_light *t, uint32_t n) { g_area = n ? t : NULL; g_n_area = t ? n : 0; }
#define IS_AREA_LIGHT(l) ((l)->type == PVO_LIGHT_SLOT || (l)->type == PV_LIGHT_AREA)
static const pvo_area_light *area_of(const pv_scene_desc *sc, const pv_light *l) {
    if (l->type == PV_LIGHT_AREA) {                                         /* an area light of the scene description itself (include/pv.h) */
        static __thread pvo_area_light view;
        view.slot = (uint32_t)(l - sc->lights); view.n_tris = l->area.n_tris; view.flags = l->area.flags; view.pad = 0;
        view.tri = sc->light_tris + 9 * (size_t)l->area.first_tri;
        memcpy(view.Lemit, l->intensity, sizeof(view.Lemit));
        return &view;
    }
    uint32_t slot = (uint32_t)(l - sc->lights);
    for (uint32_t i = 0; i < g_n_area; ++i) if (g_area[i].slot == slot) return &g_area[i];
    return NULL;
}
static float tri_area(const float *tv) {                                /* shapes/trianglemesh.cpp:284-290 */
    v3 p1 = V(tv[0], tv[1], tv[2]), p2 = V(tv[3], tv[4], tv[5]), p3 = V(tv[6], tv[7], tv[8]);
    return 0.5f * vlen(vcross(vsub(p2, p1), vsub(p3, p1)));
}
/* Triangle::Intersect's DifferentialGeometry normal for default uvs (trianglemesh.cpp:160-205, core/diffgeom.cpp:40-55) */
static v3 tri_dg_nn(const float *tv, uint32_t flags) {
    v3 p1 = V(tv[0], tv[1], tv[2]), p2 = V(tv[3], tv[4], tv[5]), p3 = V(tv[6], tv[7], tv[8]);
    float du1 = 0.f - 1.f, du2 = 1.f - 1.f, dv1 = 0.f - 1.f, dv2 = 0.f - 1.f;
    v3 dp1 = vsub(p1, p3), dp2 = vsub(p2, p3);
    float determinant = du1 * dv2 - dv1 * du2;
    float invdet = 1.f / determinant;
    v3 dpdu = vmul(vsub(vmul(dp1, dv2), vmul(dp2, dv1)), invdet);
    v3 dpdv = vmul(vadd(vmul(dp1, -du2), vmul(dp2, du1)), invdet);
    v3 nn = vnorm(vcross(dpdu, dpdv));
    if (((flags & 1u) != 0) ^ ((flags & 2u) != 0)) nn = vmul(nn, -1.f);   /* ReverseOrientation ^ TransformSwapsHandedness */
    return nn;
}
static spec area_sample_L(const pv_scene_desc *sc, const pv_light *l, v3 p, float uComp, float u0, float u1, v3 *wi, float *pdf, visray *vis) {
    const pvo_area_light *al = area_of(sc, l);
    spec L = s_const(0.f);
    *pdf = 0.f; *wi = V(0, 0, 1); vis->o = p; vis->d = *wi; vis->mint = 0.f; vis->maxt = 0.f;
    if (!al || !al->n_tris) return L;
    const uint32_t n = al->n_tris;
    /* ShapeSet's area distribution (core/light.cpp:129-136, Distribution1D core/montecarlo.h:55-107) */
    float *area = (float *)malloc(sizeof(float) * n), *cdf = (float *)malloc(sizeof(float) * (n + 1));
    float sumArea = 0.f;
    for (uint32_t i = 0; i < n; ++i) { area[i] = tri_area(al->tri + 9 * (size_t)i); sumArea += area[i]; }
    cdf[0] = 0.f;
    for (uint32_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + area[i - 1] / n;
    float funcInt = cdf[n];
    if (funcInt == 0.f) for (uint32_t i = 1; i < n + 1; ++i) cdf[i] = (float)i / (float)n;
    else for (uint32_t i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    int lo = 0, hi = (int)n + 1;                                          /* SampleDiscrete: upper_bound - 1 */
    while (lo < hi) { int mid = (lo + hi) / 2; if (uComp < cdf[mid]) hi = mid; else lo = mid + 1; }
    int sn = lo - 1; if (sn < 0) sn = 0;
    /* Triangle::Sample(u1, u2, Ns) */
    const float *tv = al->tri + 9 * (size_t)sn;
    v3 p1 = V(tv[0], tv[1], tv[2]), p2 = V(tv[3], tv[4], tv[5]), p3 = V(tv[6], tv[7], tv[8]);
    float su1 = sqrtf(u0), b1 = 1.f - su1, b2 = u1 * su1;
    v3 pt = vadd(vadd(vmul(p1, b1), vmul(p2, b2)), vmul(p3, 1.f - b1 - b2));
    v3 ns = vnorm(vcross(vsub(p2, p1), vsub(p3, p1)));
    if (al->flags & 1u) ns = vmul(ns, -1.f);
    /* ShapeSet::Sample(p, ls, Ns) (core/light.cpp:145-158): ONE ray against every shape, thit / dg of the LAST shape it hits */
    v3 rdir = vsub(pt, p);
    float thit = 1.f;
    int anyHit = 0;
    for (uint32_t i = 0; i < n; ++i) {
        float t;
        if (tri_hit(al->tri + 9 * (size_t)i, p, rdir, 1e-3f, INFINITY, &t)) { anyHit = 1; thit = t; ns = tri_dg_nn(al->tri + 9 * (size_t)i, al->flags); }
    }
    (void)anyHit;
    v3 ps = ray_at(p, rdir, thit);
    *wi = vnorm(vsub(ps, p));
    /* ShapeSet::Pdf(p, wi) (:167-172) with Shape::Pdf (core/shape.cpp:86-99) */
    float pdfsum = 0.f;
    for (uint32_t i = 0; i < n; ++i) {
        float t;
        if (!tri_hit(al->tri + 9 * (size_t)i, p, *wi, 1e-3f, INFINITY, &t)) continue;
        v3 nn = tri_dg_nn(al->tri + 9 * (size_t)i, al->flags);
        float pd = dist2(p, ray_at(p, *wi, t)) / (fabsf(vdot(nn, vneg(*wi))) * area[i]);
        if (isinf(pd)) pd = 0.f;
        pdfsum += area[i] * pd;
    }
    *pdf = pdfsum / sumArea;
    /* visibility->SetSegment(p, pEpsilon = 0, ps, 1e-3f, time) (core/light.h:85-93) */
    float dist = vlen(vsub(p, ps));
    vis->o = p; vis->d = vdiv(vsub(ps, p), dist); vis->mint = 0.f; vis->maxt = dist * (1.f - 1e-3f);
    if (vdot(ns, vneg(*wi)) > 0.f) L = s_load(al->Lemit);                  /* DiffuseAreaLight::L (lights/diffuse.h:51-53) */
    free(area); free(cdf);
    return L;
}

/* RainbowVolume::rainbowReflection volumes/rainbow.cpp:12-78 */
static float lerp_transfer(float x, float x0, float x1, float y0, float y1) {
    if (x <= x0) return y0;
    if (x >= x1) return y1;
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}
static float lerp_or_zero(float x, float x0, float x1, float y0, float y1) {
    if (x < x0 || x1 < x) return 0.f;
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}

/* ---------------------------------------------------------------- Li (integrators/photonvolume.cpp:112-222) */
typedef struct {
    int mode; mt_rng mt; uint32_t k0, k1, r0, r1;
} li_rng;

static void mt_shuffle_f(float *samp, uint32_t count, uint32_t dims, mt_rng *rng) {     /* core/montecarlo.h:175-181 */
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t other = i + (mt_next(rng) % (count - i));
        for (uint32_t j = 0; j < dims; ++j) { float t = samp[dims * i + j]; samp[dims * i + j] = samp[dims * other + j]; samp[dims * other + j] = t; }
    }
}
static float sobol2(uint32_t n, uint32_t scramble) {                                     /* core/montecarlo.h:288-293 */
    for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1) if (n & 1) scramble ^= v;
    float f = (float)((scramble >> 8) & 0xffffffu) / 16777216.0f;
    return f < ONE_MINUS_EPS ? f : ONE_MINUS_EPS;
}

static void li_one(const pv_scene_desc *sc, const pvo_kdtree *t, const float *wi_pl, const float *alpha_pl,
                   const pv_ray *ray, uint64_t ray_index, const pv_gather_params *prm, int rng_mode, uint32_t mt_seed_v,
                   closeph *buf, float *Lout, float *Tout, pv_gather_stats *st) {
    const pv_medium *vr = sc->medium;
    spec Tr = s_const(1.f), Lv = s_const(0.f);
    v3 ro = V(ray->o[0], ray->o[1], ray->o[2]), rd = V(ray->d[0], ray->d[1], ray->d[2]);
    float t0, t1;
    if (st) st->rays++;
    if (!vr || vr->type == PV_MEDIUM_NONE || !vol_intersectp(sc, ro, rd, ray->mint, ray->maxt, &t0, &t1) || (t1 - t0) == 0.f) {
        memcpy(Tout, Tr.c, sizeof(Tr.c)); memcpy(Lout, Lv.c, sizeof(Lv.c));
        return;
    }
    int rainbow = vr->type == PV_MEDIUM_RAINBOW && !g_n_more;      /* dynamic_cast<RainbowVolume*> of an AggregateVolume is NULL */
    float stepSize = prm->stepsize;
    int nSamples = (int)ceilf((t1 - t0) / stepSize);
    float step = (t1 - t0) / nSamples;
    v3 p = ray_at(ro, rd, t0), pPrev;
    v3 w = vneg(rd);
    t0 += ray->u_scatter * step;
    float maxDistSquared = prm->maxdist * prm->maxdist;
    int nLights = (int)sc->n_lights;

    li_rng rg; rg.mode = rng_mode;
    float *lightNum = NULL;
    uint32_t scramble = 0, permkey = 0;
    if (rng_mode == PVO_RNG_MT) {
        lightNum = (float *)malloc(sizeof(float) * (size_t)nSamples * 4);
    } else {
        uint32_t wds[4];
        rg.k0 = (uint32_t)prm->seed; rg.k1 = (uint32_t)(prm->seed >> 32);
        rg.r0 = (uint32_t)ray_index; rg.r1 = (uint32_t)(ray_index >> 32);
        pv_philox4x32_10(rg.r0, rg.r1, 0, PV_RNG_RAY, rg.k0, rg.k1, wds);
        scramble = wds[0]; permkey = wds[1];
    }
    if (rng_mode == PVO_RNG_MT) {
        /* LDShuffleScrambled1D(1, nSamples, ..) x2 and LDShuffleScrambled2D(1, nSamples, ..)
         * (photonvolume.cpp:137-142, montecarlo.h:304-323): nSamples(arg) = 1, nPixel = nSamples, so each
         * per-pixel Shuffle of one element still draws one RandomUInt, then the pixels are shuffled. */
        mt_seed(&rg.mt, mt_seed_v);
        float *lightComp = lightNum + nSamples, *lightPos = lightNum + 2 * (size_t)nSamples;
        uint32_t s1 = mt_next(&rg.mt);
        for (int i = 0; i < nSamples; ++i) lightNum[i] = pv_van_der_corput((uint32_t)i, s1);
        for (int i = 0; i < nSamples; ++i) mt_shuffle_f(lightNum + i, 1, 1, &rg.mt);
        mt_shuffle_f(lightNum, (uint32_t)nSamples, 1, &rg.mt);
        uint32_t s2 = mt_next(&rg.mt);
        for (int i = 0; i < nSamples; ++i) lightComp[i] = pv_van_der_corput((uint32_t)i, s2);
        for (int i = 0; i < nSamples; ++i) mt_shuffle_f(lightComp + i, 1, 1, &rg.mt);
        mt_shuffle_f(lightComp, (uint32_t)nSamples, 1, &rg.mt);
        uint32_t s3a = mt_next(&rg.mt), s3b = mt_next(&rg.mt);
        for (int i = 0; i < nSamples; ++i) { lightPos[2 * i] = pv_van_der_corput((uint32_t)i, s3a); lightPos[2 * i + 1] = sobol2((uint32_t)i, s3b); }
        for (int i = 0; i < nSamples; ++i) mt_shuffle_f(lightPos + 2 * i, 1, 2, &rg.mt);
        mt_shuffle_f(lightPos, (uint32_t)nSamples, 2, &rg.mt);
    }
    med_counters mc = {0};
    for (int i = 0; i < nSamples; ++i, t0 += step) {
        uint32_t wds[4] = {0, 0, 0, 0};
        if (rng_mode == PVO_RNG_PHILOX) pv_philox4x32_10(rg.r0, rg.r1, (uint32_t)i, PV_RNG_STEP, rg.k0, rg.k1, wds);
        pPrev = p;
        p = ray_at(ro, rd, t0);
        float u_tau = rng_mode == PVO_RNG_MT ? pv_u32_to_float(mt_next(&rg.mt)) : pv_u32_to_float(wds[0]);
        spec stepTau = vol_tau(sc, pPrev, vsub(p, pPrev), 0.f, 1.f, .5f * stepSize, u_tau, &mc);
        Tr = s_exp_neg(&stepTau);
        if (s_y(sc, &Tr) < 1e-3) {
            const float continueProb = .5f;
            float u_rr = rng_mode == PVO_RNG_MT ? pv_u32_to_float(mt_next(&rg.mt)) : pv_u32_to_float(wds[1]);
            if (u_rr > continueProb) { Tr = s_const(0.f); break; }
            for (int b = 0; b < NS; ++b) Tr.c[b] /= continueProb;
        }
        spec L_i = s_const(0.f), L_d = s_const(0.f), L_ii = s_const(0.f);
        spec ss = vol_sigma(sc, SEL_S, p, &mc);
        spec sa = vol_sigma(sc, SEL_A, p, &mc);
        if (!s_black(&ss) && nLights > 0 && !(prm->flags & PV_GATHER_NO_DIRECT)) {
            float u_l = rng_mode == PVO_RNG_MT ? lightNum[i]
                        : pv_van_der_corput(pv_permute((uint32_t)i, (uint32_t)nSamples, permkey), scramble);
            int ln = (int)floorf(u_l * nLights);
            if (ln > nLights - 1) ln = nLights - 1;
            const pv_light *light = &sc->lights[ln];
            float pdf; visray vis; v3 wo;
            spec L;
            if (IS_AREA_LIGHT(light)) {
                /* LightSample(up0, up1, ucomp) is called as ls(lightComp, lightPos[0], lightPos[1]): the component number is the THIRD
                 * argument (core/light.h).  Philox mode: words 0, 1 (position) and 2 (component) of the step's PV_RNG_AREA block */
                float uc, ua, ub;
                if (rng_mode == PVO_RNG_MT) { uc = lightNum[2 * (size_t)nSamples + 2 * i + 1]; ua = lightNum[nSamples + i]; ub = lightNum[2 * (size_t)nSamples + 2 * i]; }
                else {
                    uint32_t aw[4];
                    pv_philox4x32_10(rg.r0, rg.r1, (uint32_t)i, PV_RNG_AREA, rg.k0, rg.k1, aw);
                    uc = pv_u32_to_float(aw[2]); ua = pv_u32_to_float(aw[0]); ub = pv_u32_to_float(aw[1]);
                }
                L = area_sample_L(sc, light, p, uc, ua, ub, &wo, &pdf, &vis);
            } else L = light_sample_L_point(light, p, &wo, &pdf, &vis);
            if (!s_black(&L) && pdf > 0.f) {
                if (st) st->shadow_rays++;
                if (!bvh_intersectp(sc, vis.o, vis.d, vis.mint, vis.maxt, NULL)) {
                    /* vis.Transmittance -> PhotonVolumeIntegrator::Transmittance(sample=NULL): step 4*stepSize, offset RandomFloat */
                    float u_sh = rng_mode == PVO_RNG_MT ? pv_u32_to_float(mt_next(&rg.mt)) : pv_u32_to_float(wds[2]);
                    spec tau = vol_tau(sc, vis.o, vis.d, vis.mint, vis.maxt, 4.f * stepSize, u_sh, &mc);
                    spec Ld;
                    for (int b = 0; b < NS; ++b) Ld.c[b] = L.c[b] * expf(-tau.c[b]);
                    if (rainbow) {
                        v3 wv = rd, wiv = wo;                               /* rainbowReflection(Ld, ray.d, wo) */
                        float cosTheta = vdot(wiv, vneg(wv));
                        const float radToDeg = 57.2957f;
                        float theta = radToDeg * acosf(cosTheta);
                        float costh2 = vdot(wiv, vneg(wv));                 /* PhaseMieHazy(wi, -w) core/volume.cpp:138-141 */
                        float I = (0.5f + 4.5f * powf((float)(0.5 * (double)(1.f + costh2)), 8.f)) / (4.f * PI_F);
                        float innerGlow = lerp_transfer(theta, 40.4f, 40.45f, 1.0f, 0.9f);
                        I *= innerGlow;
                        float rainbowI = 1.0f;
                        float primaryRainbowI = 0.92f;
                        float secondaryRainbowI = (float)(0.42 * (double)primaryRainbowI);
                        float mistI = 0.08f;
                        float lambda = lerp_or_zero(theta, 40.4f, 42.3f, 400.0f, 700.0f);
                        if (lambda) rainbowI *= primaryRainbowI;
                        else {
                            lambda = lerp_or_zero(theta, 51.0f, 54.4f, 700.0f, 400.0f);
                            if (lambda) rainbowI *= secondaryRainbowI;
                        }
                        if (!lambda) {
                            for (int b = 0; b < NS; ++b) L_d.c[b] = (Ld.c[b] * (I * mistI));
                        } else {
                            /* CoefficientSpectrum::filter core/spectrum.h:300-320 */
                            spec rb = s_const(0.f);
                            float deltaLambda = (float)(700 - 400) / NS;
                            float iwd = (lambda - 400) / deltaLambda;
                            int index = (int)iwd;
                            float tt = iwd - index;
                            if (index >= 0 && index < NS) rb.c[index] = Ld.c[index] * tt;
                            if (index + 1 < NS && index + 1 >= 0) rb.c[index + 1] = Ld.c[index + 1] * (1 - tt);
                            for (int b = 0; b < NS; ++b) L_d.c[b] = (Ld.c[b] * mistI + rb.c[b] * rainbowI) * I;
                        }
                    } else {
                        float ph = vol_p(sc, p, w, vneg(wo));
                        for (int b = 0; b < NS; ++b) L_d.c[b] = ((Ld.c[b] * ph) * (float)nLights) / pdf;
                    }
                }
            }
        }
        if (!rainbow && !(prm->flags & PV_GATHER_NO_INDIRECT)) {
            spec l = lphoton(sc, t, wi_pl, alpha_pl, w, p, prm->nused, maxDistSquared, buf, st);
            for (int b = 0; b < NS; ++b) L_ii.c[b] += l.c[b];
        }
        if (s_y(sc, &sa) != 0.0 || s_y(sc, &ss) != 0.0) {
            for (int b = 0; b < NS; ++b) L_i.c[b] = L_d.c[b] + (ss.c[b] / (sa.c[b] + ss.c[b])) * L_ii.c[b];
        } else L_i = L_d;
        /* Lve */
        spec lve = vol_sigma(sc, SEL_LE, p, &mc);
        for (int b = 0; b < NS; ++b)
            Lv.c[b] = ((sa.c[b] * lve.c[b]) * step) + ((ss.c[b] * L_i.c[b]) * step) + (Tr.c[b] * Lv.c[b]);
    }
    if (st) st->density_samples += mc.density_samples;
    free(lightNum);
    memcpy(Tout, Tr.c, sizeof(Tr.c)); memcpy(Lout, Lv.c, sizeof(Lv.c));
}

typedef struct {
    const pv_scene_desc *sc; const pvo_kdtree *t; const float *wi, *alpha; const pv_ray *rays; uint64_t n;
    const pv_gather_params *prm; int rng_mode; uint32_t mt_seed; float *L, *T; pv_gather_stats st;
    uint64_t *next; pthread_mutex_t *mu;
} gather_job;
static void *gather_worker(void *arg) {
    gather_job *j = (gather_job *)arg;
    closeph *buf = (closeph *)malloc(sizeof(closeph) * (j->prm->nused ? j->prm->nused : 1));
    const uint64_t chunk = 256;
    for (;;) {
        pthread_mutex_lock(j->mu);
        uint64_t b = *j->next; *j->next += chunk;
        pthread_mutex_unlock(j->mu);
        if (b >= j->n) break;
        uint64_t e = b + chunk < j->n ? b + chunk : j->n;
        for (uint64_t i = b; i < e; ++i)
            li_one(j->sc, j->t, j->wi, j->alpha, &j->rays[i], j->prm->ray_index_base + i, j->prm, j->rng_mode,
                   j->mt_seed + (uint32_t)i, buf, j->L + NS * i, j->T + NS * i, &j->st);
    }
    free(buf);
    return NULL;
}
int pvo_gather(const pv_scene_desc *sc, const pvo_kdtree *t, const float *wi, const float *alpha, const pv_ray *rays,
               uint64_t n, const pv_gather_params *prm, int rng_mode, uint32_t mt_seed_v, int nthreads, float *L, float *T,
               pv_gather_stats *stats) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    uint64_t next = 0;
    gather_job jobs[256]; pthread_t th[256];
    for (int k = 0; k < nthreads; ++k) {
        gather_job j = {sc, t, wi, alpha, rays, n, prm, rng_mode, mt_seed_v, L, T, {0}, &next, &mu};
        jobs[k] = j;
    }
    if (nthreads == 1) gather_worker(&jobs[0]);
    else {
        for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, gather_worker, &jobs[k]);
        for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        for (int k = 0; k < nthreads; ++k) {
            stats->rays += jobs[k].st.rays; stats->lookups += jobs[k].st.lookups;
            stats->photons_found += jobs[k].st.photons_found; stats->heap_lookups += jobs[k].st.heap_lookups;
            stats->shadow_rays += jobs[k].st.shadow_rays; stats->density_samples += jobs[k].st.density_samples;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- the other two volume integrators
 * SingleScatteringIntegrator::Li  integrators/single.cpp:66-138
 * EmissionIntegrator::Li          integrators/emission.cpp:63-106
 * Both march like PhotonVolumeIntegrator::Li but keep a CUMULATIVE transmittance (Tr *= Exp(-stepTau), single.cpp:100,
 * emission.cpp:89), test the Russian roulette on it, add Tr * Lve(p) per step and -- single only -- the light's
 * single-scattered radiance Tr * ss * p(p, w, -wo) * Ld * nLights / pdf (single.cpp:128-129); the sum is scaled by the
 * step length at the end.  MT mode replays the reference's draw order (three LDShuffleScrambled tables first for
 * "single", none for "emission"; then per step: tau offset, roulette if taken, shadow-ray tau offset), Philox mode uses
 * the streams of the CUDA path (the same counters as pvo_gather). */
static void vli_one(const pv_scene_desc *sc, const pv_ray *ray, uint64_t ray_index, const pv_gather_params *prm, int kind,
                    int rng_mode, uint32_t mt_seed_v, float *Lout, float *Tout, pv_gather_stats *st) {
    const pv_medium *vr = sc->medium;
    spec Tr = s_const(1.f), Lv = s_const(0.f);
    v3 ro = V(ray->o[0], ray->o[1], ray->o[2]), rd = V(ray->d[0], ray->d[1], ray->d[2]);
    float t0, t1;
    if (st) st->rays++;
    if (!vr || vr->type == PV_MEDIUM_NONE || !vol_intersectp(sc, ro, rd, ray->mint, ray->maxt, &t0, &t1) || (t1 - t0) == 0.f) {
        memcpy(Tout, Tr.c, sizeof(Tr.c)); memcpy(Lout, Lv.c, sizeof(Lv.c));
        return;
    }
    const int single = kind == PVO_VLI_SINGLE;
    float stepSize = prm->stepsize;
    int nSamples = (int)ceilf((t1 - t0) / stepSize);
    float step = (t1 - t0) / nSamples;
    v3 p = ray_at(ro, rd, t0), pPrev;
    v3 w = vneg(rd);
    t0 += ray->u_scatter * step;
    int nLights = (int)sc->n_lights;

    li_rng rg; rg.mode = rng_mode;
    float *lightNum = NULL;
    uint32_t scramble = 0, permkey = 0;
    if (rng_mode == PVO_RNG_MT) {
        mt_seed(&rg.mt, mt_seed_v);
        if (single) {                                           /* single.cpp:85-90, montecarlo.h:304-323 */
            lightNum = (float *)malloc(sizeof(float) * (size_t)nSamples * 4);
            float *lightComp = lightNum + nSamples, *lightPos = lightNum + 2 * (size_t)nSamples;
            uint32_t s1 = mt_next(&rg.mt);
            for (int i = 0; i < nSamples; ++i) lightNum[i] = pv_van_der_corput((uint32_t)i, s1);
            for (int i = 0; i < nSamples; ++i) mt_shuffle_f(lightNum + i, 1, 1, &rg.mt);
            mt_shuffle_f(lightNum, (uint32_t)nSamples, 1, &rg.mt);
            uint32_t s2 = mt_next(&rg.mt);
            for (int i = 0; i < nSamples; ++i) lightComp[i] = pv_van_der_corput((uint32_t)i, s2);
            for (int i = 0; i < nSamples; ++i) mt_shuffle_f(lightComp + i, 1, 1, &rg.mt);
            mt_shuffle_f(lightComp, (uint32_t)nSamples, 1, &rg.mt);
            uint32_t s3a = mt_next(&rg.mt), s3b = mt_next(&rg.mt);
            for (int i = 0; i < nSamples; ++i) { lightPos[2 * i] = pv_van_der_corput((uint32_t)i, s3a); lightPos[2 * i + 1] = sobol2((uint32_t)i, s3b); }
            for (int i = 0; i < nSamples; ++i) mt_shuffle_f(lightPos + 2 * i, 1, 2, &rg.mt);
            mt_shuffle_f(lightPos, (uint32_t)nSamples, 2, &rg.mt);
        }
    } else {
        uint32_t wds[4];
        rg.k0 = (uint32_t)prm->seed; rg.k1 = (uint32_t)(prm->seed >> 32);
        rg.r0 = (uint32_t)ray_index; rg.r1 = (uint32_t)(ray_index >> 32);
        pv_philox4x32_10(rg.r0, rg.r1, 0, PV_RNG_RAY, rg.k0, rg.k1, wds);
        scramble = wds[0]; permkey = wds[1];
    }
    med_counters mc = {0};
    for (int i = 0; i < nSamples; ++i, t0 += step) {
        uint32_t wds[4] = {0, 0, 0, 0};
        if (rng_mode == PVO_RNG_PHILOX) pv_philox4x32_10(rg.r0, rg.r1, (uint32_t)i, PV_RNG_STEP, rg.k0, rg.k1, wds);
        pPrev = p;
        p = ray_at(ro, rd, t0);
        float u_tau = rng_mode == PVO_RNG_MT ? pv_u32_to_float(mt_next(&rg.mt)) : pv_u32_to_float(wds[0]);
        spec stepTau = vol_tau(sc, pPrev, vsub(p, pPrev), 0.f, 1.f, .5f * stepSize, u_tau, &mc);
        for (int b = 0; b < NS; ++b) Tr.c[b] *= expf(-stepTau.c[b]);
        if (s_y(sc, &Tr) < 1e-3) {
            const float continueProb = .5f;
            float u_rr = rng_mode == PVO_RNG_MT ? pv_u32_to_float(mt_next(&rg.mt)) : pv_u32_to_float(wds[1]);
            if (u_rr > continueProb) { Tr = s_const(0.f); break; }
            for (int b = 0; b < NS; ++b) Tr.c[b] /= continueProb;
        }
        spec lve = vol_sigma(sc, SEL_LE, p, &mc);
        for (int b = 0; b < NS; ++b) Lv.c[b] += Tr.c[b] * lve.c[b];
        if (!single) continue;
        spec ss = vol_sigma(sc, SEL_S, p, &mc);
        if (!s_black(&ss) && nLights > 0) {
            float u_l = rng_mode == PVO_RNG_MT ? lightNum[i]
                        : pv_van_der_corput(pv_permute((uint32_t)i, (uint32_t)nSamples, permkey), scramble);
            int ln = (int)floorf(u_l * nLights);
            if (ln > nLights - 1) ln = nLights - 1;
            const pv_light *light = &sc->lights[ln];
            float pdf; visray vis; v3 wo;
            spec L;
            if (IS_AREA_LIGHT(light)) {
                /* LightSample(up0, up1, ucomp) is called as ls(lightComp, lightPos[0], lightPos[1]): the component number is the THIRD
                 * argument (core/light.h).  Philox mode: words 0, 1 (position) and 2 (component) of the step's PV_RNG_AREA block */
                float uc, ua, ub;
                if (rng_mode == PVO_RNG_MT) { uc = lightNum[2 * (size_t)nSamples + 2 * i + 1]; ua = lightNum[nSamples + i]; ub = lightNum[2 * (size_t)nSamples + 2 * i]; }
                else {
                    uint32_t aw[4];
                    pv_philox4x32_10(rg.r0, rg.r1, (uint32_t)i, PV_RNG_AREA, rg.k0, rg.k1, aw);
                    uc = pv_u32_to_float(aw[2]); ua = pv_u32_to_float(aw[0]); ub = pv_u32_to_float(aw[1]);
                }
                L = area_sample_L(sc, light, p, uc, ua, ub, &wo, &pdf, &vis);
            } else L = light_sample_L_point(light, p, &wo, &pdf, &vis);
            if (!s_black(&L) && pdf > 0.f) {
                if (st) st->shadow_rays++;
                if (!bvh_intersectp(sc, vis.o, vis.d, vis.mint, vis.maxt, NULL)) {
                    /* vis.Transmittance -> SingleScatteringIntegrator::Transmittance(sample = NULL): step 4 * stepSize (single.cpp:55-58) */
                    float u_sh = rng_mode == PVO_RNG_MT ? pv_u32_to_float(mt_next(&rg.mt)) : pv_u32_to_float(wds[2]);
                    spec tau = vol_tau(sc, vis.o, vis.d, vis.mint, vis.maxt, 4.f * stepSize, u_sh, &mc);
                    float ph = vol_p(sc, p, w, vneg(wo));
                    for (int b = 0; b < NS; ++b) {
                        float Ld = L.c[b] * expf(-tau.c[b]);
                        Lv.c[b] += ((((Tr.c[b] * ss.c[b]) * ph) * Ld) * (float)nLights) / pdf;
                    }
                }
            }
        }
    }
    if (st) st->density_samples += mc.density_samples;
    free(lightNum);
    for (int b = 0; b < NS; ++b) Lv.c[b] *= step;
    memcpy(Tout, Tr.c, sizeof(Tr.c)); memcpy(Lout, Lv.c, sizeof(Lv.c));
}
int pvo_volume_li(const pv_scene_desc *sc, const pv_ray *rays, uint64_t n, const pv_gather_params *prm, int kind, int rng_mode,
                  uint32_t mt_seed_v, float *L, float *T, pv_gather_stats *stats) {
    if (kind != PVO_VLI_SINGLE && kind != PVO_VLI_EMISSION) return -1;
    pv_gather_stats st; memset(&st, 0, sizeof(st));
    for (uint64_t i = 0; i < n; ++i)
        vli_one(sc, &rays[i], prm->ray_index_base + i, prm, kind, rng_mode, mt_seed_v + (uint32_t)i, L + NS * i, T + NS * i, &st);
    if (stats) *stats = st;
    return 0;
}

/* ---------------------------------------------------------------- photon shooting (core/photonshooter.cpp:47-357) */
typedef struct {
    uint64_t n, cap;
    float *pos, *wi, *alpha; uint64_t *ids;
} phvec;
static void phvec_push(phvec *v, v3 p, const spec *a, v3 w, uint64_t id) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 4096;
        v->pos = (float *)realloc(v->pos, sizeof(float) * 3 * v->cap);
        v->wi = (float *)realloc(v->wi, sizeof(float) * 3 * v->cap);
        v->alpha = (float *)realloc(v->alpha, sizeof(float) * NS * v->cap);
        v->ids = (uint64_t *)realloc(v->ids, sizeof(uint64_t) * v->cap);
    }
    v->pos[3 * v->n] = p.x; v->pos[3 * v->n + 1] = p.y; v->pos[3 * v->n + 2] = p.z;
    v->wi[3 * v->n] = w.x; v->wi[3 * v->n + 1] = w.y; v->wi[3 * v->n + 2] = w.z;
    memcpy(v->alpha + NS * v->n, a->c, sizeof(a->c));
    v->ids[v->n] = id;
    v->n++;
}

typedef struct {
    const pv_scene_desc *sc;
    const pv_shoot_params *prm;
    stream *rng;
    phvec *local;
    uint64_t path_index; uint32_t deposit_seq;
    bvh_counters bc; med_counters mc; uint64_t segments;
    /* surface maps (photonshooter.cpp:147-189).  All zero == the volume-only pass: causticDone and indirectDone true,
       volumeDone false -- no surface deposit, no radiance photon, diffuse bounces end the path (Q6). */
    phvec *surf[4];                 /* caustic, indirect, direct, radiance sites (wi := n, alpha := rho_r) */
    int want_caustic, want_indirect, volume_done, final_gather;
    uint64_t first_hit_scatters;    /* shooter->nVolumePaths++ (:104) */
} shoot_ctx;
#define PVO_ID(cls, path, seq) (((uint64_t)(cls) << 60) | ((uint64_t)(path) << 16) | ((uint64_t)(seq) & 0xffffu))
/* skip n draws of the path stream (BSDF::rho's stratified samples are drawn and, for a Lambertian BRDF, never used) */
static void st_skip(stream *s, uint32_t n) {
    if (s->mode == PVO_RNG_MT) { for (uint32_t i = 0; i < n; ++i) (void)mt_next(s->mt); return; }
    uint64_t consumed = (uint64_t)s->j * 4 - (4 - s->pos) + n;
    uint32_t q = (uint32_t)(consumed / 4), r = (uint32_t)(consumed % 4);
    if (r == 0) { s->j = q; s->pos = 4; }
    else { pv_philox4x32_10(s->c0, s->c1, q, PV_RNG_PATH, s->k0, s->k1, s->buf); s->j = q + 1; s->pos = r; }
}

typedef struct { int prim; v3 p; v3 nn, dpdu; float rayEpsilon; } isect_t;
typedef struct { v3 o, d; float mint, maxt; } ray_t;

/* renderer->Transmittance(scene, ray, NULL, rng) -> photonvolume.cpp:15-30 with sample == NULL */
static spec shoot_transmittance(shoot_ctx *c, const ray_t *r) {
    float offset = st_float(c->rng);
    spec tau = vol_tau(c->sc, r->o, r->d, r->mint, r->maxt, 4.f * c->prm->integrator_stepsize, offset, &c->mc);
    return s_exp_neg(&tau);
}

/* dg + shading frame for a triangle hit: shapes/trianglemesh.cpp:160-205, core/diffgeom.cpp:40-55 */
static void make_isect(const pv_scene_desc *sc, int prim, v3 o, v3 d, float t, isect_t *is) {
    if (sc->n_spheres && sc->prim_shape[prim] != PV_SHAPE_TRIANGLE) {
        /* shapes/sphere.cpp:112-163 + core/diffgeom.cpp:40-55 */
        const pv_sphere *s = &sc->spheres[sc->prim_shape[prim]];
        v3 ro = xf_point(s->world_to_object, o), rd = xf_vec(s->world_to_object, d);
        v3 phit; float phi;
        sphere_phit(s, ro, rd, t, &phit, &phi);
        float cz = phit.z / s->radius; cz = cz < -1.f ? -1.f : (cz > 1.f ? 1.f : cz);
        float theta = acosf(cz);
        float zradius = sqrtf(phit.x * phit.x + phit.y * phit.y);
        float invzradius = 1.f / zradius;
        float cosphi = phit.x * invzradius, sinphi = phit.y * invzradius;
        v3 dpdu = V(-s->phi_max * phit.y, s->phi_max * phit.x, 0);
        v3 dpdv = vmul(V(phit.z * cosphi, phit.z * sinphi, -s->radius * sinf(theta)), s->theta_max - s->theta_min);
        v3 wdpdu = xf_vec(s->object_to_world, dpdu), wdpdv = xf_vec(s->object_to_world, dpdv);
        is->prim = prim;
        is->p = xf_point(s->object_to_world, phit);
        is->dpdu = wdpdu;
        is->nn = vnorm(vcross(wdpdu, wdpdv));
        if (s->flip_normal) is->nn = vmul(is->nn, -1.f);
        is->rayEpsilon = 5e-4f * t;
        return;
    }
    const float *tv = sc->tri_verts + 9 * (size_t)prim;
    v3 p1 = V(tv[0], tv[1], tv[2]), p2 = V(tv[3], tv[4], tv[5]), p3 = V(tv[6], tv[7], tv[8]);
    /* default uvs (0,0),(1,0),(1,1): du1=-1 du2=0 dv1=-1 dv2=-1, determinant 1 */
    float du1 = 0.f - 1.f, du2 = 1.f - 1.f, dv1 = 0.f - 1.f, dv2 = 0.f - 1.f;
    v3 dp1 = vsub(p1, p3), dp2 = vsub(p2, p3);
    float determinant = du1 * dv2 - dv1 * du2;
    float invdet = 1.f / determinant;
    v3 dpdu = vmul(vsub(vmul(dp1, dv2), vmul(dp2, dv1)), invdet);
    v3 dpdv = vmul(vadd(vmul(dp1, -du2), vmul(dp2, du1)), invdet);
    is->prim = prim;
    is->p = ray_at(o, d, t);
    is->dpdu = dpdu;
    is->nn = vnorm(vcross(dpdu, dpdv));
    is->rayEpsilon = 1e-3f * t;
}

/* split_lambda: the fork's Spectrum::lambda member as followPhoton sees it (photonshooter.cpp:141).  It is STATE, not a function of
 * the bins at the time of the test: every Spectrum produced by a binary operator is a CoefficientSpectrum converted back through
 * SampledSpectrum(const CoefficientSpectrum&), which sets lambda = extractLambda() (core/spectrum.h:339-343: 400 + 10 i when exactly
 * one bin is positive, else -1); compound assignments keep it.  So lambda is recomputed where alpha is re-made -- at emission
 * (`alpha = AbsDot * Le / pdf`, :263) and after every surface bounce (`alpha = anew / continueProb`, :217) -- and merely carried
 * through the in-place updates in between (`alpha *= ref; alpha /= pdf` at a scatter event :119-120, `alpha *= Transmittance` :133).
 * A monochromatic child whose one bin underflows to zero in a dense medium therefore reaches the next glass face with lambda >= 0:
 * no split, one BSDF sample, and (0/0 -> continueProb 1) the black photon is traced on -- now with lambda = -1, so that the NEXT
 * dispersive face splits it into nothing and ends the path. */
static int s_extract_lambda(const spec *a) {             /* CoefficientSpectrum::extractLambda, core/spectrum.h:266-279 */
    int l = -1, first = 1;
    for (int b = 0; b < NS; ++b) {
        if (a->c[b] > 0.f && !first) return -1;
        if (a->c[b] > 0.f && first) { l = 400 + b * 10; first = 0; }
    }
    return l;
}
static void follow_photon(shoot_ctx *c, ray_t photonRay, isect_t photonIsect, spec alpha, int nIntersections, int specularPath, int split_lambda) {
    const pv_scene_desc *sc = c->sc;
    float thit = photonRay.maxt;
    c->segments++;
    int prim = bvh_intersect(sc, photonRay.o, photonRay.d, photonRay.mint, &thit, &c->bc);
    if (prim < 0) return;
    make_isect(sc, prim, photonRay.o, photonRay.d, thit, &photonIsect);
    photonRay.maxt = thit;                                  /* GeometricPrimitive::Intersect sets r.maxt (primitive.cpp:172) */
    ++nIntersections;
    float t0, t1;
    float length = vlen(photonRay.d);
    if (length == 0.f) return;
    v3 rnd = vdiv(photonRay.d, length);
    float rn_mint = photonRay.mint * length, rn_maxt = photonRay.maxt * length;
    if (!vol_intersectp(sc, photonRay.o, rnd, rn_mint, rn_maxt, &t0, &t1)) { t0 = 1.0f; t1 = 0.0f; }
    t0 += st_float(c->rng) * c->prm->stepsize;
    float t_i = t0;
    float xi = st_float(c->rng);
    int interaction = 0;
    while (t0 < t1) {
        ray_t shortRay = {photonRay.o, rnd, t_i, t0};
        spec tr = shoot_transmittance(c, &shortRay);
        if (xi > s_y(sc, &tr)) { interaction = 1; break; }
        t0 += c->prm->stepsize;
    }
    if (interaction) {
        v3 interactPt = ray_at(photonRay.o, rnd, t0);
        spec sig_s = vol_sigma(sc, SEL_S, interactPt, &c->mc);
        spec sig_a = vol_sigma(sc, SEL_A, interactPt, &c->mc);
        float ys = s_y(sc, &sig_s), ya = s_y(sc, &sig_a);
        int scatter = (st_float(c->rng) > (ys) / (ya + ys));     /* Q1: inverted test, photonshooter.cpp:88 */
        if (!scatter) return;
        if (!c->volume_done) {                                   /* `if (scatter && !volumeDone)` :96 */
        if (nIntersections > 1) {
            uint64_t id = PVO_ID(0, c->path_index, c->deposit_seq);
            c->deposit_seq++;
            phvec_push(c->local, interactPt, &alpha, rnd, id);
        } else c->first_hit_scatters++;
        float u1 = st_float(c->rng);
        float u2 = st_float(c->rng);
        v3 direction = uniform_sample_sphere(u1, u2);
        float pdf = 1.f / (4.f * PI_F);
        float ref = vol_p(sc, interactPt, rnd, direction);
        if (ref == 0.f || pdf == 0.f) return;
        for (int b = 0; b < NS; ++b) alpha.c[b] *= ref;
        for (int b = 0; b < NS; ++b) alpha.c[b] /= pdf;
        photonRay.o = interactPt; photonRay.d = direction; photonRay.mint = 0.f; photonRay.maxt = INFINITY;
        follow_photon(c, photonRay, photonIsect, alpha, nIntersections, specularPath, split_lambda);
        /* Q2: falls through into the surface code with the scattered ray and the ORIGINAL isect */
        }
    }
    {
        spec tr = shoot_transmittance(c, &photonRay);
        for (int b = 0; b < NS; ++b) alpha.c[b] *= tr.c[b];
    }
    const pv_material *mat = &sc->materials[sc->prim_material[photonIsect.prim]];
    v3 wo = vneg(photonRay.d);
    /* surface deposits (photonshooter.cpp:147-189); hasNonSpecular == a matte surface with a non-black Kd (materials/matte.cpp:55) */
    {
        spec Kd = s_load(mat->kd);
        int hasNonSpecular = mat->type == PV_MAT_MATTE && !s_black(&Kd);
        if (hasNonSpecular) {
            int cls = -1;
            if (specularPath && nIntersections > 1) { if (c->want_caustic) cls = 0; }
            else if (nIntersections == 1 && c->want_indirect && c->final_gather) cls = 2;
            else if (nIntersections > 1 && c->want_indirect) cls = 1;
            if (cls >= 0) {
                phvec_push(c->surf[cls], photonIsect.p, &alpha, wo, PVO_ID(cls + 1, c->path_index, c->deposit_seq));
                c->deposit_seq++;
                if (c->final_gather && st_float(c->rng) < .125f) {
                    v3 n = photonIsect.nn;
                    if (vdot(n, vneg(photonRay.d)) < 0.f) n = vneg(n);          /* Faceforward(n, -photonRay.d) */
                    /* rho_r = BSDF::rho(rng, BSDF_ALL_REFLECTION) == Kd for a Lambertian (reflection.cpp:647-659, :326);
                       rho_t == 0; each rho draws 2 x StratifiedSample2D(6x6) = 144 floats */
                    phvec_push(c->surf[3], photonIsect.p, &Kd, n, PVO_ID(4, c->path_index, c->deposit_seq));
                    c->deposit_seq++;
                    st_skip(c->rng, 288);
                }
            }
        }
    }
    if (nIntersections >= c->prm->max_photon_depth) return;
    /* BSDF frame core/reflection.cpp:619-627 */
    v3 nn = photonIsect.nn, ng = photonIsect.nn;
    v3 sn = vnorm(photonIsect.dpdu);
    v3 tn = vcross(nn, sn);
    if (mat->type == PV_MAT_MATTE) {
        float uDir0 = st_float(c->rng), uDir1 = st_float(c->rng), uComp = st_float(c->rng);
        (void)uComp;
        spec R = s_load(mat->kd);
        if (s_black(&R)) return;                                 /* no BxDF: Sample_f returns 0 (reflection.cpp:541-546) */
        v3 wol = V(vdot(wo, sn), vdot(wo, tn), vdot(wo, nn));
        v3 wil;
        concentric_sample_disk(uDir0, uDir1, &wil.x, &wil.y);
        wil.z = sqrtf(fmaxf(0.f, 1.f - wil.x * wil.x - wil.y * wil.y));
        if (wol.z < 0.f) wil.z *= -1.f;
        float pdf = (wol.z * wil.z > 0.f) ? fabsf(wil.z) * INV_PI_F : 0.f;
        if (pdf == 0.f) return;
        v3 wiW = V(sn.x * wil.x + tn.x * wil.y + nn.x * wil.z, sn.y * wil.x + tn.y * wil.y + nn.y * wil.z,
                   sn.z * wil.x + tn.z * wil.y + nn.z * wil.z);
        int reflect = vdot(wiW, ng) * vdot(wo, ng) > 0;
        if (!reflect) return;                                    /* f == 0 -> fr.IsBlack() */
        spec fr; for (int b = 0; b < NS; ++b) fr.c[b] = R.c[b] * INV_PI_F;
        float ad = fabsf(vdot(wiW, nn));
        spec anew; for (int b = 0; b < NS; ++b) anew.c[b] = ((alpha.c[b] * fr.c[b]) * ad) / pdf;
        float continueProb = fminf(1.f, s_y(sc, &anew) / s_y(sc, &alpha));
        if (st_float(c->rng) > continueProb) return;
        /* specularPath &= false; indirectDone && !specularPath -> continue (Q6) */
        if (!c->want_indirect) return;
        spec an2; for (int b = 0; b < NS; ++b) an2.c[b] = anew.c[b] / continueProb;
        ray_t nr = {photonIsect.p, wiW, photonIsect.rayEpsilon, INFINITY};
        follow_photon(c, nr, photonIsect, an2, nIntersections, 0, s_extract_lambda(&an2));
        return;
    }
    /* glass: specular reflection + dispersive transmission (materials/glass.cpp:42-59,
       core/reflection.cpp:115-182, core/spectrum.h:253-279) */
    {
        spec Rk = s_load(mat->kr), Tk = s_load(mat->kt);
        int hasR = !s_black(&Rk), hasT = !s_black(&Tk);
        int matching = hasR + hasT;
        /* split (photonshooter.cpp:140-145): hasTransmission && alpha.lambda < 0 && primitive->dispersive() */
        int nspec = 1; int binlist[NS];
        int do_split = hasT && split_lambda < 0 && mat->vn > 0.f;
        if (do_split) { nspec = 0; for (int b = 0; b < NS; ++b) if (alpha.c[b] != 0.f) binlist[nspec++] = b; }
        spec alpha_in = alpha;
        for (int si = 0; si < nspec; ++si) {
            spec a = alpha_in;
            if (do_split) { a = s_const(0.f); a.c[binlist[si]] = alpha_in.c[binlist[si]]; }
            float uDir0 = st_float(c->rng), uDir1 = st_float(c->rng), uComp = st_float(c->rng);
            (void)uDir0; (void)uDir1;
            if (matching == 0) continue;
            int which = (int)floorf(uComp * matching);
            if (which > matching - 1) which = matching - 1;
            int pickT = hasR ? (which == 1) : 1;
            v3 wol = V(vdot(wo, sn), vdot(wo, tn), vdot(wo, nn));
            v3 wil; spec f; float pdf = 1.f;
            float ior = mat->index;
            if (!pickT) {
                wil = V(-wol.x, -wol.y, wol.z);
                /* FresnelDielectric(1, ior).Evaluate(CosTheta(wo)) */
                float cosi = wol.z < -1.f ? -1.f : (wol.z > 1.f ? 1.f : wol.z);
                int entering = cosi > 0.;
                float ei = 1.f, et = ior;
                if (!entering) { float tt = ei; ei = et; et = tt; }
                float sint = ei / et * sqrtf(fmaxf(0.f, 1.f - cosi * cosi));
                float F;
                if (sint >= 1.) F = 1.f;
                else {
                    float cost = sqrtf(fmaxf(0.f, 1.f - sint * sint));
                    float ac = fabsf(cosi);
                    float Rparl = ((et * ac) - (ei * cost)) / ((et * ac) + (ei * cost));
                    float Rperp = ((ei * ac) - (et * cost)) / ((ei * ac) + (et * cost));
                    F = (Rparl * Rparl + Rperp * Rperp) / 2.f;
                }
                for (int b = 0; b < NS; ++b) f.c[b] = (F * Rk.c[b]) / fabsf(wil.z);
            } else {
                int entering = wol.z > 0.;
                float ei = 1.f, et = ior;
                /* extractLambda (spectrum.h:266-279): integer step (700-400)/(30-1) == 10 */
                int lam = -1, firstb = 1;
                for (int b = 0; b < NS; ++b) { if (a.c[b] > 0.f && !firstb) { lam = -1; break; } if (a.c[b] > 0.f && firstb) { lam = 400 + b * 10; firstb = 0; } }
                if (lam > 0 && mat->vn > 0.f) {
                    float l = lam / 1000.f;
                    float B = (float)(((et - 1) / mat->vn) * 0.52345);
                    float A = (float)(et - (B / 0.34522792));
                    et = (float)(A + B / pow((double)l, 2));
                }
                float ei0 = ei, et0 = et;
                if (!entering) { float tt = ei; ei = et; et = tt; }
                float sini2 = fmaxf(0.f, 1.f - wol.z * wol.z);
                float eta = ei / et;
                float sint2 = eta * eta * sini2;
                if (sint2 >= 1.) continue;                      /* TIR: returns 0 with pdf untouched (0) */
                float cost = sqrtf(fmaxf(0.f, 1.f - sint2));
                if (entering) cost = -cost;
                wil = V(eta * -wol.x, eta * -wol.y, cost);
                /* fresnel member is FresnelDielectric(etai, etat) with the UNdispersed etat */
                float cosi = wol.z < -1.f ? -1.f : (wol.z > 1.f ? 1.f : wol.z);
                float fei = 1.f, fet = ior; (void)ei0; (void)et0;
                if (!(cosi > 0.)) { float tt = fei; fei = fet; fet = tt; }
                float sint = fei / fet * sqrtf(fmaxf(0.f, 1.f - cosi * cosi));
                float F;
                if (sint >= 1.) F = 1.f;
                else {
                    float cost2 = sqrtf(fmaxf(0.f, 1.f - sint * sint));
                    float ac = fabsf(cosi);
                    float Rparl = ((fet * ac) - (fei * cost2)) / ((fet * ac) + (fei * cost2));
                    float Rperp = ((fei * ac) - (fet * cost2)) / ((fei * ac) + (fet * cost2));
                    F = (Rparl * Rparl + Rperp * Rperp) / 2.f;
                }
                for (int b = 0; b < NS; ++b) f.c[b] = ((1.f - F) * Tk.c[b]) / fabsf(wil.z);
            }
            if (matching > 1) pdf /= matching;
            if (s_black(&f) || pdf == 0.f) continue;
            v3 wiW = V(sn.x * wil.x + tn.x * wil.y + nn.x * wil.z, sn.y * wil.x + tn.y * wil.y + nn.y * wil.z,
                       sn.z * wil.x + tn.z * wil.y + nn.z * wil.z);
            float ad = fabsf(vdot(wiW, nn));
            spec anew; for (int b = 0; b < NS; ++b) anew.c[b] = ((a.c[b] * f.c[b]) * ad) / pdf;
            float continueProb = fminf(1.f, s_y(sc, &anew) / s_y(sc, &a));
            if (st_float(c->rng) > continueProb) continue;
            spec an2; for (int b = 0; b < NS; ++b) an2.c[b] = anew.c[b] / continueProb;
            /* specular: specularPath stays as it was; `indirectDone && !specularPath` -> continue */
            if (!c->want_indirect && !specularPath) continue;
            ray_t nr = {photonIsect.p, wiW, photonIsect.rayEpsilon, INFINITY};
            follow_photon(c, nr, photonIsect, an2, nIntersections, specularPath, s_extract_lambda(&an2));
        }
    }
}

/* Light::Sample_L(scene, ls, u1, u2, time, &ray, &Ns, &pdf): lights/point.cpp:80-88, spot.cpp:106-114, distant.cpp:82-102 */
/* DiffuseAreaLight::Sample_L(scene, ls, u1, u2, ...) (lights/diffuse.cpp:89-100): a point by area (ShapeSet::Sample(ls, Ns),
 * core/light.cpp:161-164), a direction uniform over the sphere flipped into the normal's hemisphere, pdf = ShapeSet::Pdf(org) / 2 pi
 * where ShapeSet::Pdf(p) = sum_i areas[i] * (1 / areas[i]) / sumArea (:175-180, core/shape.cpp:80-82) -- the NUMBER of shapes over the
 * total area, not 1 / area. */
static spec area_emit(const pv_scene_desc *sc, const pv_light *l, float up0, float up1, float ucomp, float u1, float u2, ray_t *ray, v3 *Ns, float *pdf) {
    const pvo_area_light *al = area_of(sc, l);
    spec Le = s_const(0.f);
    ray->o = V(0, 0, 0); ray->d = V(0, 0, 1); ray->mint = 1e-3f; ray->maxt = INFINITY; *Ns = ray->d; *pdf = 0.f;
    if (!al || !al->n_tris) return Le;
    const uint32_t n = al->n_tris;
    float *area = (float *)malloc(sizeof(float) * n), *cdf = (float *)malloc(sizeof(float) * (n + 1));
    float sumArea = 0.f;
    for (uint32_t i = 0; i < n; ++i) { area[i] = tri_area(al->tri + 9 * (size_t)i); sumArea += area[i]; }
    cdf[0] = 0.f;
    for (uint32_t i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + area[i - 1] / n;
    float funcInt = cdf[n];
    if (funcInt == 0.f) for (uint32_t i = 1; i < n + 1; ++i) cdf[i] = (float)i / (float)n;
    else for (uint32_t i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
    int lo = 0, hi = (int)n + 1;
    while (lo < hi) { int mid = (lo + hi) / 2; if (ucomp < cdf[mid]) hi = mid; else lo = mid + 1; }
    int sn = lo - 1; if (sn < 0) sn = 0;
    const float *tv = al->tri + 9 * (size_t)sn;
    v3 p1 = V(tv[0], tv[1], tv[2]), p2 = V(tv[3], tv[4], tv[5]), p3 = V(tv[6], tv[7], tv[8]);
    float su1 = sqrtf(up0), b1 = 1.f - su1, b2 = up1 * su1;
    v3 org = vadd(vadd(vmul(p1, b1), vmul(p2, b2)), vmul(p3, 1.f - b1 - b2));
    v3 ns = vnorm(vcross(vsub(p2, p1), vsub(p3, p1)));
    if (al->flags & 1u) ns = vmul(ns, -1.f);
    v3 dir = uniform_sample_sphere(u1, u2);
    if (vdot(dir, ns) < 0.) dir = vmul(dir, -1.f);
    float pd = 0.f;
    for (uint32_t i = 0; i < n; ++i) pd += area[i] * (1.f / area[i]);
    *pdf = (pd / sumArea) * 0.15915494309189533577f;                    /* INV_TWOPI */
    ray->o = org; ray->d = dir; *Ns = ns;
    if (vdot(ns, dir) > 0.f) Le = s_load(al->Lemit);
    free(area); free(cdf);
    return Le;
}
static spec light_emit(const pv_scene_desc *sc, const pv_light *l, float up0, float up1, ray_t *ray, v3 *Ns, float *pdf) {
    spec Le = s_load(l->intensity);
    if (l->type == PV_LIGHT_POINT) {
        ray->o = V(l->pos[0], l->pos[1], l->pos[2]); ray->d = uniform_sample_sphere(up0, up1);
        *pdf = 1.f / (4.f * PI_F);
    } else if (l->type == PV_LIGHT_SPOT) {
        v3 v = uniform_sample_cone(up0, up1, l->cos_total_width);
        ray->o = V(l->pos[0], l->pos[1], l->pos[2]); ray->d = xf_vec(l->light_to_world, v);
        *pdf = 1.f / (2.f * PI_F * (1.f - l->cos_total_width));
        float fo = spot_falloff(l, ray->d);
        for (int b = 0; b < NS; ++b) Le.c[b] = Le.c[b] * fo;
    } else {
        const float *wb = sc->world_bound;
        v3 pmin = V(wb[0], wb[1], wb[2]), pmax = V(wb[3], wb[4], wb[5]);
        v3 wc = vadd(vmul(pmin, .5f), vmul(pmax, .5f));
        float wr = bbox_inside(wb, wb + 3, wc) ? vlen(vsub(wc, pmax)) : 0.f;
        v3 ld = V(l->dir[0], l->dir[1], l->dir[2]), v1, v2;
        coordinate_system(ld, &v1, &v2);
        float d1, d2;
        concentric_sample_disk(up0, up1, &d1, &d2);
        v3 Pdisk = vadd(wc, vmul(vadd(vmul(v1, d1), vmul(v2, d2)), wr));
        ray->o = vadd(Pdisk, vmul(ld, wr)); ray->d = vneg(ld);
        *pdf = 1.f / (PI_F * wr * wr);
    }
    ray->mint = 0.f; ray->maxt = INFINITY;
    *Ns = ray->d;
    return Le;
}

typedef struct { float *func, *cdf; float funcInt; int count; } distrib1d;
static void distrib_init(distrib1d *d, const pv_scene_desc *sc) {          /* core/montecarlo.h:55-83 */
    int n = (int)sc->n_lights;
    d->count = n; d->func = (float *)malloc(sizeof(float) * n); d->cdf = (float *)malloc(sizeof(float) * (n + 1));
    for (int i = 0; i < n; ++i) d->func[i] = sc->lights[i].power_y;
    d->cdf[0] = 0.f;
    for (int i = 1; i < n + 1; ++i) d->cdf[i] = d->cdf[i - 1] + d->func[i - 1] / n;
    d->funcInt = d->cdf[n];
    if (d->funcInt == 0.f) for (int i = 1; i < n + 1; ++i) d->cdf[i] = (float)i / (float)n;
    else for (int i = 1; i < n + 1; ++i) d->cdf[i] /= d->funcInt;
}
static int distrib_sample_discrete(const distrib1d *d, float u, float *pdf) {  /* :99-107 */
    int lo = 0, hi = d->count + 1;            /* std::upper_bound(cdf, cdf+count+1, u) */
    while (lo < hi) { int mid = (lo + hi) / 2; if (u < d->cdf[mid]) hi = mid; else lo = mid + 1; }
    int offset = lo - 1; if (offset < 0) offset = 0;
    *pdf = d->func[offset] / (d->funcInt * d->count);
    return offset;
}

/* one light path = body of the for-loop photonshooter.cpp:248-277 */
static void shoot_path(shoot_ctx *c, const halton6 *h, const distrib1d *ld, uint64_t path_index) {
    const pv_scene_desc *sc = c->sc;
    float u[6];
    halton6_sample(h, (uint32_t)path_index, u);
    float lightPdf;
    int lightNum = distrib_sample_discrete(ld, u[0], &lightPdf);
    const pv_light *light = &sc->lights[lightNum];
    ray_t photonRay; v3 Nl; float pdf;
    spec Le = IS_AREA_LIGHT(light) ? area_emit(sc, light, u[1], u[2], u[3], u[4], u[5], &photonRay, &Nl, &pdf)
                                            : light_emit(sc, light, u[1], u[2], &photonRay, &Nl, &pdf);
    if (pdf == 0.f || s_black(&Le)) return;
    float ad = fabsf(vdot(Nl, photonRay.d));
    spec alpha; float den = pdf * lightPdf;
    for (int b = 0; b < NS; ++b) alpha.c[b] = (Le.c[b] * ad) / den;
    if (s_black(&alpha)) return;
    isect_t is; memset(&is, 0, sizeof(is));
    c->path_index = path_index; c->deposit_seq = 0;
    follow_photon(c, photonRay, is, alpha, 0, 1, s_extract_lambda(&alpha));
}

typedef struct {
    const pv_scene_desc *sc; const pv_shoot_params *prm; const halton6 *h; const distrib1d *ld;
    uint64_t *next_block; uint64_t end_block; pthread_mutex_t *mu;
    phvec *per_block;      /* indexed by block - first_block */
    uint64_t first_block;
    bvh_counters bc; med_counters mc; uint64_t segments;
} shoot_job;
static void *shoot_worker(void *arg) {
    shoot_job *j = (shoot_job *)arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        uint64_t b = *j->next_block; *j->next_block += 1;
        pthread_mutex_unlock(j->mu);
        if (b >= j->end_block) break;
        stream st; memset(&st, 0, sizeof(st));
        st.mode = PVO_RNG_PHILOX; st.k0 = (uint32_t)j->prm->seed; st.k1 = (uint32_t)(j->prm->seed >> 32);
        shoot_ctx c; memset(&c, 0, sizeof(c));
        c.sc = j->sc; c.prm = j->prm; c.rng = &st; c.local = &j->per_block[b - j->first_block];
        for (uint64_t i = 0; i < 4096; ++i) {
            uint64_t path = (b - 1) * 4096 + i + 1;
            st.c0 = (uint32_t)path; st.c1 = (uint32_t)(path >> 32); st.j = 0; st.pos = 4;
            shoot_path(&c, j->h, j->ld, path);
        }
        j->bc.nodes_visited += c.bc.nodes_visited; j->bc.tri_tests += c.bc.tri_tests;
        j->mc.density_samples += c.mc.density_samples; j->segments += c.segments;
    }
    return NULL;
}

static void merge_block(pvo_photons *out, phvec *acc, phvec *blk, uint64_t nshot) {
    /* photonshooter.cpp:330-340: alpha /= float(nshot) at merge time */
    float fn = (float)(uint32_t)nshot;
    for (uint64_t i = 0; i < blk->n; ++i) {
        spec a = s_load(blk->alpha + NS * i);
        for (int b = 0; b < NS; ++b) a.c[b] /= fn;
        phvec_push(acc, V(blk->pos[3 * i], blk->pos[3 * i + 1], blk->pos[3 * i + 2]), &a,
                   V(blk->wi[3 * i], blk->wi[3 * i + 1], blk->wi[3 * i + 2]), blk->ids[i]);
    }
    (void)out;
}
static void phvec_free(phvec *v) { free(v->pos); free(v->wi); free(v->alpha); free(v->ids); memset(v, 0, sizeof(*v)); }

int pvo_shoot(const pv_scene_desc *sc, uint64_t n_wanted, const pv_shoot_params *prm, int rng_mode, int nthreads,
              pvo_photons *out) {
    memset(out, 0, sizeof(*out));
    if (!sc->n_lights || !sc->medium) return PV_EINVAL;
    mt_rng mt; mt_seed(&mt, 0u);                       /* RNG rng(31 * taskNum), taskNum == 0 */
    halton6 h; halton6_init(&h, &mt);                  /* PermutedHalton halton(6, rng) */
    distrib1d ld; distrib_init(&ld, sc);
    phvec acc; memset(&acc, 0, sizeof(acc));
    uint64_t nshot = 0, block = 0;
    uint64_t max_paths = prm->max_paths ? prm->max_paths : ((uint64_t)1 << 40);
    int rc = 0;
    bvh_counters bc = {0, 0}; med_counters mc = {0}; uint64_t segments = 0;
    if (rng_mode == PVO_RNG_MT) {
        stream st; memset(&st, 0, sizeof(st)); st.mode = PVO_RNG_MT; st.mt = &mt;
        phvec local; memset(&local, 0, sizeof(local));
        shoot_ctx c; memset(&c, 0, sizeof(c));
        c.sc = sc; c.prm = prm; c.rng = &st; c.local = &local;
        uint32_t totalPaths = 0;
        for (;;) {
            for (uint32_t i = 0; i < 4096; ++i) shoot_path(&c, &h, &ld, ++totalPaths);
            block++;
            if (nshot > 500000 && acc.n < n_wanted && (acc.n == 0 || acc.n < 4096 / 1024)) { rc = PV_ENOPHOTONS; acc.n = 0; break; }
            nshot += 4096;
            merge_block(out, &acc, &local, nshot);
            local.n = 0;
            if (acc.n >= n_wanted) break;
            if (nshot >= max_paths) break;
        }
        phvec_free(&local);
        bc = c.bc; mc = c.mc; segments = c.segments;
    } else {
        if (nthreads < 1) nthreads = 1;
        if (nthreads > 256) nthreads = 256;
        uint64_t wave = (uint64_t)nthreads * 4;
        pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
        int done = 0;
        while (!done) {
            uint64_t first = block + 1, end = first + wave, next = first;
            phvec *per = (phvec *)calloc(wave, sizeof(phvec));
            shoot_job jobs[256]; pthread_t th[256];
            for (int k = 0; k < nthreads; ++k) {
                shoot_job j = {sc, prm, &h, &ld, &next, end, &mu, per, first, {0, 0}, {0}, 0};
                jobs[k] = j;
            }
            if (nthreads == 1) shoot_worker(&jobs[0]);
            else {
                for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, shoot_worker, &jobs[k]);
                for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
            }
            for (int k = 0; k < nthreads; ++k) {
                bc.nodes_visited += jobs[k].bc.nodes_visited; bc.tri_tests += jobs[k].bc.tri_tests;
                mc.density_samples += jobs[k].mc.density_samples; segments += jobs[k].segments;
            }
            for (uint64_t b = 0; b < wave && !done; ++b) {
                block++;
                if (nshot > 500000 && acc.n < n_wanted && (acc.n == 0 || acc.n < 4096 / 1024)) { rc = PV_ENOPHOTONS; acc.n = 0; done = 1; break; }
                nshot += 4096;
                merge_block(out, &acc, &per[b], nshot);
                if (acc.n >= n_wanted || nshot >= max_paths) done = 1;
            }
            for (uint64_t b = 0; b < wave; ++b) phvec_free(&per[b]);
            free(per);
        }
    }
    free(ld.func); free(ld.cdf);
    out->n = acc.n; out->pos = acc.pos; out->wi = acc.wi; out->alpha = acc.alpha; out->ids = acc.ids;
    out->nshot = nshot; out->blocks = block;
    out->nodes_visited = bc.nodes_visited; out->tri_tests = bc.tri_tests; out->density_samples = mc.density_samples;
    out->segments = segments;
    return rc;
}
/* PhotonShootingTask::Run with all maps (photonshooter.cpp:232-357), ONE task: per-class done flags that change at block
 * ends (:303-341), per-class path counts, the give-up rule over the three wanted counts (:285-299). */
int pvo_shoot_maps(const pv_scene_desc *sc, uint64_t n_volume, uint64_t n_caustic, uint64_t n_indirect, int final_gather,
                   const pv_shoot_params *prm, int rng_mode, pvo_maps *out) {
    memset(out, 0, sizeof(*out));
    if (!sc->n_lights || !sc->medium) return PV_EINVAL;
    mt_rng mt; mt_seed(&mt, 0u);
    halton6 h; halton6_init(&h, &mt);
    distrib1d ld; distrib_init(&ld, sc);
    phvec acc[5], local[5]; memset(acc, 0, sizeof(acc)); memset(local, 0, sizeof(local));
    stream st; memset(&st, 0, sizeof(st)); st.mode = rng_mode; st.mt = &mt;
    st.k0 = (uint32_t)prm->seed; st.k1 = (uint32_t)(prm->seed >> 32);
    shoot_ctx c; memset(&c, 0, sizeof(c));
    c.sc = sc; c.prm = prm; c.rng = &st; c.local = &local[0];
    for (int k = 0; k < 4; ++k) c.surf[k] = &local[k + 1];
    c.final_gather = final_gather;
    int causticDone = n_caustic == 0, indirectDone = n_indirect == 0, volumeDone = n_volume == 0;
    uint64_t nshot = 0, path = 0;
    uint64_t max_paths = prm->max_paths ? prm->max_paths : ((uint64_t)1 << 40);
    int rc = 0;
    for (;;) {
        c.want_caustic = !causticDone; c.want_indirect = !indirectDone; c.volume_done = volumeDone;
        for (uint32_t i = 0; i < 4096; ++i) {
            ++path;
            if (rng_mode != PVO_RNG_MT) { st.c0 = (uint32_t)path; st.c1 = (uint32_t)(path >> 32); st.j = 0; st.pos = 4; }
            shoot_path(&c, &h, &ld, path);
        }
        out->blocks++;
#define UNSUCC(needed, found) ((found) < (needed) && ((found) == 0 || (found) < 4096 / 1024))
        if (nshot > 500000 && (UNSUCC(n_caustic, acc[1].n) || UNSUCC(n_indirect, acc[2].n) || UNSUCC(n_volume, acc[0].n))) {
            rc = PV_ENOPHOTONS; acc[0].n = acc[1].n = acc[2].n = acc[4].n = 0; break;       /* :292-296 (direct photons are kept) */
        }
#undef UNSUCC
        nshot += 4096;
        if (!indirectDone) {
            out->n_indirect_paths += 4096;
            for (uint64_t i = 0; i < local[2].n; ++i) { spec a = s_load(local[2].alpha + NS * i);
                phvec_push(&acc[2], V(local[2].pos[3*i], local[2].pos[3*i+1], local[2].pos[3*i+2]), &a, V(local[2].wi[3*i], local[2].wi[3*i+1], local[2].wi[3*i+2]), local[2].ids[i]); }
            if (acc[2].n >= n_indirect) indirectDone = 1;
            out->n_direct_paths += 4096;
            for (uint64_t i = 0; i < local[3].n; ++i) { spec a = s_load(local[3].alpha + NS * i);
                phvec_push(&acc[3], V(local[3].pos[3*i], local[3].pos[3*i+1], local[3].pos[3*i+2]), &a, V(local[3].wi[3*i], local[3].wi[3*i+1], local[3].wi[3*i+2]), local[3].ids[i]); }
        }
        local[2].n = local[3].n = 0;
        if (!causticDone) {
            out->n_caustic_paths += 4096;
            for (uint64_t i = 0; i < local[1].n; ++i) { spec a = s_load(local[1].alpha + NS * i);
                phvec_push(&acc[1], V(local[1].pos[3*i], local[1].pos[3*i+1], local[1].pos[3*i+2]), &a, V(local[1].wi[3*i], local[1].wi[3*i+1], local[1].wi[3*i+2]), local[1].ids[i]); }
            if (acc[1].n >= n_caustic) causticDone = 1;
        }
        local[1].n = 0;
        if (!volumeDone) {
            out->n_volume_paths += 4096;
            merge_block(NULL, &acc[0], &local[0], nshot);
            if (acc[0].n >= n_volume) volumeDone = 1;
        }
        local[0].n = 0;
        for (uint64_t i = 0; i < local[4].n; ++i) { spec a = s_load(local[4].alpha + NS * i);
            phvec_push(&acc[4], V(local[4].pos[3*i], local[4].pos[3*i+1], local[4].pos[3*i+2]), &a, V(local[4].wi[3*i], local[4].wi[3*i+1], local[4].wi[3*i+2]), local[4].ids[i]); }
        local[4].n = 0;
        if (indirectDone && causticDone && volumeDone) break;
        if (nshot >= max_paths) break;
    }
    out->n_volume_paths += c.first_hit_scatters;
    for (int k = 0; k < 5; ++k) {
        phvec_free(&local[k]);
        out->cls[k].n = acc[k].n; out->cls[k].pos = acc[k].pos; out->cls[k].wi = acc[k].wi; out->cls[k].alpha = acc[k].alpha; out->cls[k].ids = acc[k].ids;
        out->cls[k].nshot = nshot; out->cls[k].blocks = out->blocks;
    }
    out->nshot = nshot;
    free(ld.func); free(ld.cdf);
    return rc;
}
void pvo_maps_free(pvo_maps *m) { for (int k = 0; k < 5; ++k) pvo_photons_free(&m->cls[k]); }

/* EPhoton + ComputeRadianceTask::Run (photonshooter.cpp:17-35,359-395) for radiance photons with rho_t == 0:
 * Lo = INV_PI * rho_r * (E_direct + E_indirect + E_caustic), E = sum of alpha over the found photons with n.wi > 0,
 * divided by count * md2 * pi, md2 = the search radius^2 as the lookup left it. */
static void ephoton_acc(const pvo_kdtree *t, const float *wi, const float *alpha, uint64_t count, uint32_t nLookup, float maxDist2,
                        v3 p, v3 n, closeph *buf, spec *E) {
    if (!t || t->nNodes == 0) return;
    photon_proc proc = {buf, nLookup, 0, 0};
    float md2 = maxDist2;
    kd_lookup(t, 0, p, &proc, &md2);
    if (proc.nFound == 0) return;
    spec e = s_const(0.f);
    for (uint32_t i = 0; i < proc.nFound; ++i) {
        uint32_t o = t->nodeOrig[buf[i].node];
        if (vdot(n, V(wi[3 * o], wi[3 * o + 1], wi[3 * o + 2])) > 0.f)
            for (int b = 0; b < NS; ++b) e.c[b] += alpha[NS * (size_t)o + b];
    }
    float den = (float)((double)((float)(int)count * md2) * 3.14159265358979323846);     /* count * md2 * M_PI: float * double */
    for (int b = 0; b < NS; ++b) E->c[b] += e.c[b] / den;
}
int pvo_radiance(const pvo_kdtree *maps[3], const float *wis[3], const float *alphas[3], const uint64_t counts[3],
                 const float *rp_pos, const float *rp_n, const float *rho_r, uint64_t n, uint32_t nLookup, float maxDist2, float *Lo) {
    closeph *buf = (closeph *)malloc(sizeof(closeph) * (nLookup ? nLookup : 1));
    for (uint64_t i = 0; i < n; ++i) {
        v3 p = V(rp_pos[3 * i], rp_pos[3 * i + 1], rp_pos[3 * i + 2]), nn = V(rp_n[3 * i], rp_n[3 * i + 1], rp_n[3 * i + 2]);
        spec r = s_load(rho_r + NS * i), E = s_const(0.f);
        if (!s_black(&r)) for (int k = 0; k < 3; ++k) ephoton_acc(maps[k], wis[k], alphas[k], counts[k], nLookup, maxDist2, p, nn, buf, &E);
        for (int b = 0; b < NS; ++b) Lo[NS * i + b] = (INV_PI_F * r.c[b]) * E.c[b];
    }
    free(buf);
    return 0;
}

/* PhotonIntegrator's LPhoton, diffuse branch (integrators/photonmap.cpp:62-108; kernel() :57-60): Lr / Lt before the
 * rho * INV_PI factors. */
int pvo_surface_lphoton(const pvo_kdtree *t, const float *wi, const float *alpha, const float *pts, const float *nf, uint64_t n,
                        uint32_t nLookup, float maxDist2, uint64_t nPaths, float *Lr, float *Lt) {
    closeph *buf = (closeph *)malloc(sizeof(closeph) * (nLookup ? nLookup : 1));
    for (uint64_t i = 0; i < n; ++i) {
        spec lr = s_const(0.f), lt = s_const(0.f);
        if (t && t->nNodes) {
            v3 p = V(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), Nf = V(nf[3 * i], nf[3 * i + 1], nf[3 * i + 2]);
            photon_proc proc = {buf, nLookup, 0, 0};
            float md2 = maxDist2;
            kd_lookup(t, 0, p, &proc, &md2);
            for (uint32_t k = 0; k < proc.nFound; ++k) {
                uint32_t o = t->nodeOrig[buf[k].node];
                float s = (1.f - dist2(t->nodePos[buf[k].node], p) / md2);
                float kern = 3.f * INV_PI_F * s * s;
                float w = kern / ((int)nPaths * md2);
                spec *dst = vdot(Nf, V(wi[3 * o], wi[3 * o + 1], wi[3 * o + 2])) > 0.f ? &lr : &lt;
                for (int b = 0; b < NS; ++b) dst->c[b] += w * alpha[NS * (size_t)o + b];
            }
        }
        memcpy(Lr + NS * i, lr.c, sizeof(lr.c)); memcpy(Lt + NS * i, lt.c, sizeof(lt.c));
    }
    free(buf);
    return 0;
}
/* RadiancePhotonProcess + KdTree::Lookup(p, proc, INFINITY) (core/photonshooter.h:54-70): the nearest radiance photon whose
 * normal faces the query normal.  Stated as the rule the CUDA path implements: smallest (d2, index) among the facing photons
 * (the reference's kd-tree keeps the first one it visits among exact ties). */
int pvo_radiance_nearest(const float *rp_pos, const float *rp_n, uint64_t n_rp, const float *pts, const float *nrm, uint64_t n,
                         uint32_t *idx, float *d2out) {
    for (uint64_t i = 0; i < n; ++i) {
        v3 p = V(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), nn = V(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
        float best = INFINITY; uint32_t bi = 0xFFFFFFFFu;
        for (uint64_t j = 0; j < n_rp; ++j) {
            if (!(vdot(V(rp_n[3 * j], rp_n[3 * j + 1], rp_n[3 * j + 2]), nn) > 0.f)) continue;
            float d2 = dist2(V(rp_pos[3 * j], rp_pos[3 * j + 1], rp_pos[3 * j + 2]), p);
            if (d2 < best) { best = d2; bi = (uint32_t)j; }
        }
        idx[i] = bi; if (d2out) d2out[i] = best;
    }
    return 0;
}

/* One batch of final-gather rays (integrators/photonmap.cpp:231-243): Scene::Intersect, Faceforward(hit normal, -d), nearest
 * facing radiance photon, Lo * renderer->Transmittance(ray, NULL) with the offset drawn from the keyed Philox stream. */
int pvo_final_gather(const pv_scene_desc *sc, const float *rp_pos, const float *rp_n, const float *rp_Lo, uint64_t n_rp, const pv_ray *rays,
                     uint64_t n, float step, uint64_t seed, uint64_t index_base, float *Lindir, uint32_t *idx) {
    for (uint64_t i = 0; i < n; ++i) {
        const pv_ray *r = &rays[i];
        v3 o = V(r->o[0], r->o[1], r->o[2]), d = V(r->d[0], r->d[1], r->d[2]);
        float thit = r->maxt;
        int prim = bvh_intersect(sc, o, d, r->mint, &thit, NULL);
        uint32_t bi = 0xFFFFFFFFu;
        spec L = s_const(0.f);
        if (prim >= 0) {
            isect_t is; make_isect(sc, prim, o, d, thit, &is);
            v3 nn = is.nn;
            if (vdot(nn, vneg(d)) < 0.f) nn = vneg(nn);
            float q[3] = {is.p.x, is.p.y, is.p.z}, qn[3] = {nn.x, nn.y, nn.z};
            pvo_radiance_nearest(rp_pos, rp_n, n_rp, q, qn, 1, &bi, NULL);
            if (bi != 0xFFFFFFFFu) {
                spec T = s_const(1.f);
                if (sc->medium && sc->medium->type != PV_MEDIUM_NONE) {
                    uint32_t w[4]; uint64_t ri = index_base + i;
                    pv_philox4x32_10((uint32_t)ri, (uint32_t)(ri >> 32), 0u, PV_RNG_FINAL_GATHER, (uint32_t)seed, (uint32_t)(seed >> 32), w);
                    spec tau = vol_tau(sc, o, d, r->mint, thit, step, pv_u32_to_float(w[0]), NULL);
                    T = s_exp_neg(&tau);
                }
                for (int b = 0; b < NS; ++b) L.c[b] = rp_Lo[NS * (size_t)bi + b] * T.c[b];
            }
        }
        memcpy(Lindir + NS * i, L.c, sizeof(L.c));
        if (idx) idx[i] = bi;
    }
    return 0;
}

void pvo_photons_free(pvo_photons *p) {
    free(p->pos); free(p->wi); free(p->alpha); free(p->ids);
    memset(p, 0, sizeof(*p));
}
