// pv_device.cuh -- device-side value math of the photon-volume path (sm_100a).
//
// Restates, in the reference's operation order and with unfused fp32 arithmetic
// (the whole library is compiled with -fmad=false; the reference's x86-64 build
// has no FMA), the value code the hot path calls:
//   core/geometry.h/.cpp (Vector ops, Cross in double, BBox::IntersectP/Inside),
//   core/transform.h, volumes/{homogeneous,volumegrid}, core/volume.cpp (PhaseHG,
//   DensityRegion::tau), accelerators/bvh.cpp traversal, shapes/trianglemesh.cpp,
//   lights/{point,spot,distant}.cpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/pv.h"
#include "../../include/pv_rng.h"

#define PV_PI_F 3.14159265358979323846f
#define PV_INV_PI_F 0.31830988618379067154f
#define PV_ONE_MINUS_EPS 0.99999994f
#define PV_MAX_LIGHTS 16
#define PV_FULL 0xffffffffu

struct DevMedium {
    static constexpr bool kExp = true;      // may be an exponential medium (see MedViewPlain)
    int type;
    float w2v[16];
    float p0[3], p1[3];
    float sigma_a[PV_NSPEC], sigma_s[PV_NSPEC], le[PV_NSPEC];
    float g;
    int nx, ny, nz;
    const float *density;
    int identity;          // world_to_volume is exactly the identity
};

struct DevScene {
    const pv_bvh_node *nodes; uint32_t n_nodes;
    const float *tri; const uint32_t *prim_mat; uint32_t n_prims;
    const pv_material *mats; uint32_t n_mats;
    const pv_light *lights; uint32_t n_lights;
    DevMedium med;
    float world_bound[6];
    float cie_y[PV_NSPEC];
    float light_func[PV_MAX_LIGHTS], light_cdf[PV_MAX_LIGHTS + 1], light_func_int;
    // primitives that are spheres (shapes/sphere.cpp).  On the device a sphere primitive is TAGGED IN THE TRIANGLE TABLE:
    // the first float of its 9-float slot is a NaN whose low 22 bits are the index into spheres[] (pv_set_scene writes it),
    // so the leaf loop pays one compare on a value it loads anyway and triangle-only scenes load nothing extra.
    const pv_sphere *spheres; uint32_t n_spheres;
    const uint8_t *mat_flags;          // per material: PV_MATF_* (which of Kd / Kr / Kt has a non-zero bin), made once by pv_set_scene
    // area lights (PV_LIGHT_AREA): the triangles of every ShapeSet, their areas, and per light the normalised area CDF
    // (n_tris + 1 entries at ltri_cdf + lcdf_off[light]), the total area and ShapeSet::Pdf(Point)'s numerator -- all computed once
    // by pv_set_scene with the reference's own float operations (core/light.cpp:114-137, core/montecarlo.h:55-83)
    const float *ltris, *ltri_area, *ltri_cdf;
    float larea_sum[PV_MAX_LIGHTS], larea_pd[PV_MAX_LIGHTS];
    uint32_t lcdf_off[PV_MAX_LIGHTS];
};
#define PV_MATF_KD 1
#define PV_MATF_KR 2
#define PV_MATF_KT 4
#define PV_SPHERE_TAG 0x7FC00000u
#define PV_SPHERE_INDEX_MASK 0x003FFFFFu

// The geometric part of a DevMedium copied into registers at kernel start.  The scene lives in global memory behind a
// pointer the compiler must assume the kernel's own stores may alias, so every density tap would otherwise re-load the
// extent, the grid dimensions and the grid pointer.  Same field names as DevMedium: the medium functions below are
// templates over either.
struct MedView {
    static constexpr bool kExp = true;
    int type, identity, nx, ny, nz;
    float g;
    float p0[3], p1[3];
    const float *density;
    float w2v[16];
};
// the same view for kernels instantiated for scenes WITHOUT an exponential medium: the density sampler then carries no
// trace of that case (the shooter is sensitive to every instruction in it)
struct MedViewPlain : MedView { static constexpr bool kExp = false; };
template <class MV>
__device__ __forceinline__ MV make_medview(const DevMedium &m) {
    MV v;
    v.type = m.type; v.identity = m.identity; v.nx = m.nx; v.ny = m.ny; v.nz = m.nz; v.g = m.g; v.density = m.density;
#pragma unroll
    for (int i = 0; i < 3; ++i) { v.p0[i] = m.p0[i]; v.p1[i] = m.p1[i]; }
#pragma unroll
    for (int i = 0; i < 16; ++i) v.w2v[i] = m.w2v[i];
    return v;
}

struct v3 { float x, y, z; };
__device__ __forceinline__ v3 V3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 operator-(v3 a) { return V3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ v3 operator*(v3 a, float f) { return V3(f * a.x, f * a.y, f * a.z); }
// core/geometry.h:92-96: Vector / f multiplies by the reciprocal (IEEE division, no fast math)
__device__ __forceinline__ v3 vdiv(v3 a, float f) { float inv = __fdiv_rn(1.f, f); return V3(a.x * inv, a.y * inv, a.z * inv); }
__device__ __forceinline__ float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float vlen(v3 a) { return __fsqrt_rn(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ v3 vnorm(v3 a) { return vdiv(a, vlen(a)); }
// core/geometry.h:477-484: operands widened to double, rounded once
__device__ __forceinline__ v3 vcross(v3 a, v3 b) {
    double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return V3((float)(__dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by))),
              (float)(__dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz))),
              (float)(__dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx))));
}
__device__ __forceinline__ v3 ray_at(v3 o, v3 d, float t) { return o + d * t; }
__device__ __forceinline__ float vcomp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__device__ __forceinline__ float dist2(v3 a, v3 b) { v3 d = a - b; return d.x * d.x + d.y * d.y + d.z * d.z; }

__device__ __forceinline__ v3 xf_point(const float *m, v3 p) {          // core/transform.h:192-201
    float x = p.x, y = p.y, z = p.z;
    float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
    float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
    float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
    float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
    if (wp == 1.f) return V3(xp, yp, zp);
    float inv = __fdiv_rn(1.f, wp);
    return V3(inv * xp, inv * yp, inv * zp);
}
__device__ __forceinline__ v3 xf_vec(const float *m, v3 v) {            // core/transform.h:215-221
    float x = v.x, y = v.y, z = v.z;
    return V3(m[0] * x + m[1] * y + m[2] * z, m[4] * x + m[5] * y + m[6] * z, m[8] * x + m[9] * y + m[10] * z);
}
template <class Med>
__device__ __forceinline__ v3 med_to_volume_p(const Med &m, v3 p) { return m.identity ? p : xf_point(m.w2v, p); }
template <class Med>
__device__ __forceinline__ v3 med_to_volume_v(const Med &m, v3 v) { return m.identity ? v : xf_vec(m.w2v, v); }

// core/geometry.cpp:68-86
__device__ __forceinline__ bool bbox_intersectp(const float *p0, const float *p1, v3 o, v3 d, float mint, float maxt,
                                                float *ht0, float *ht1) {
    float t0 = mint, t1 = maxt;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float invRayDir = __fdiv_rn(1.f, vcomp(d, i));
        float tNear = (p0[i] - vcomp(o, i)) * invRayDir;
        float tFar = (p1[i] - vcomp(o, i)) * invRayDir;
        if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; }
        t0 = tNear > t0 ? tNear : t0;
        t1 = tFar < t1 ? tFar : t1;
        if (t0 > t1) return false;
    }
    *ht0 = t0; *ht1 = t1;
    return true;
}
__device__ __forceinline__ bool bbox_inside(const float *p0, const float *p1, v3 p) {   // core/geometry.h:404-408
    return p.x >= p0[0] && p.x <= p1[0] && p.y >= p0[1] && p.y <= p1[1] && p.z >= p0[2] && p.z <= p1[2];
}

template <class Med>
__device__ __forceinline__ bool med_is_homog(const Med &m) { return m.type == PV_MEDIUM_HOMOGENEOUS || m.type == PV_MEDIUM_RAINBOW; }
template <class Med>
__device__ __forceinline__ bool med_intersectp(const Med &m, v3 o, v3 d, float mint, float maxt, float *t0, float *t1) {
    return bbox_intersectp(m.p0, m.p1, med_to_volume_p(m, o), med_to_volume_v(m, d), mint, maxt, t0, t1);
}
__device__ __forceinline__ float lerpf(float t, float a, float b) { return (1.f - t) * a + t * b; }   // core/pbrt.h:218
template <class Med>
__device__ __forceinline__ float grid_D(const Med &m, int x, int y, int z) {                    // volumes/volumegrid.h:60-65
    x = min(max(x, 0), m.nx - 1); y = min(max(y, 0), m.ny - 1); z = min(max(z, 0), m.nz - 1);
    return __ldg(m.density + ((size_t)z * m.nx * m.ny + (size_t)y * m.nx + x));
}
// ExponentialDensity::Density (volumes/exponential.h:57-61) for a point inside the extent, given Pobj - extent.pMin;
// e = {a, b, updir.xyz}.  Out of line: the trilinear sampler below is inlined in several hot loops and must stay small.
static __device__ __noinline__ float exp_density(const float *e, float dx, float dy, float dz) {
    const float height = dx * __ldg(e + 2) + dy * __ldg(e + 3) + dz * __ldg(e + 4);
    return __ldg(e) * expf(-__ldg(e + 1) * height);
}
// volumes/volumegrid.cpp:39-57 (and the exponential medium, which shares the DensityRegion code paths)
template <class Med>
__device__ __forceinline__ float grid_density(const Med &m, v3 Pobj) {
    if (!bbox_inside(m.p0, m.p1, Pobj)) return 0.f;
    if (Med::kExp && m.type == PV_MEDIUM_EXPONENTIAL) return exp_density(m.density, Pobj.x - m.p0[0], Pobj.y - m.p0[1], Pobj.z - m.p0[2]);
    float vx_ = __fdiv_rn(Pobj.x - m.p0[0], m.p1[0] - m.p0[0]);
    float vy_ = __fdiv_rn(Pobj.y - m.p0[1], m.p1[1] - m.p0[1]);
    float vz_ = __fdiv_rn(Pobj.z - m.p0[2], m.p1[2] - m.p0[2]);
    vx_ = vx_ * m.nx - .5f; vy_ = vy_ * m.ny - .5f; vz_ = vz_ * m.nz - .5f;
    int vx = (int)floorf(vx_), vy = (int)floorf(vy_), vz = (int)floorf(vz_);
    float dx = vx_ - vx, dy = vy_ - vy, dz = vz_ - vz;
#ifndef PV_TAPS_SHARED_CLAMP
#define PV_TAPS_SHARED_CLAMP 1
#endif
#if PV_TAPS_SHARED_CLAMP
    // the eight D(x, y, z) taps (volumegrid.h:60-65) with each coordinate clamped ONCE and 32-bit offsets (the API rejects grids
    // of 2^31 voxels or more): the same eight values as eight grid_D calls for a third of the integer work
    const int x0 = min(max(vx, 0), m.nx - 1), x1 = min(max(vx + 1, 0), m.nx - 1);
    const int y0 = min(max(vy, 0), m.ny - 1) * m.nx, y1 = min(max(vy + 1, 0), m.ny - 1) * m.nx;
    const int sl = m.nx * m.ny;
    const float *z0 = m.density + (size_t)(min(max(vz, 0), m.nz - 1) * sl), *z1 = m.density + (size_t)(min(max(vz + 1, 0), m.nz - 1) * sl);
    float d00 = lerpf(dx, __ldg(z0 + (y0 + x0)), __ldg(z0 + (y0 + x1)));
    float d10 = lerpf(dx, __ldg(z0 + (y1 + x0)), __ldg(z0 + (y1 + x1)));
    float d01 = lerpf(dx, __ldg(z1 + (y0 + x0)), __ldg(z1 + (y0 + x1)));
    float d11 = lerpf(dx, __ldg(z1 + (y1 + x0)), __ldg(z1 + (y1 + x1)));
#else
    float d00 = lerpf(dx, grid_D(m, vx, vy, vz), grid_D(m, vx + 1, vy, vz));
    float d10 = lerpf(dx, grid_D(m, vx, vy + 1, vz), grid_D(m, vx + 1, vy + 1, vz));
    float d01 = lerpf(dx, grid_D(m, vx, vy, vz + 1), grid_D(m, vx + 1, vy, vz + 1));
    float d11 = lerpf(dx, grid_D(m, vx, vy + 1, vz + 1), grid_D(m, vx + 1, vy + 1, vz + 1));
#endif
    float d0 = lerpf(dy, d00, d10);
    float d1 = lerpf(dy, d01, d11);
    return lerpf(dz, d0, d1);
}
// scalar "density" multiplying the constant spectra: 1/0 inside/outside for homogeneous media
// (volumes/homogeneous.h:62-70), the trilinear density for grids (core/volume.h:80-88)
template <class Med>
__device__ __forceinline__ float med_density(const Med &m, v3 p, uint32_t *nsamples) {
    v3 po = med_to_volume_p(m, p);
    if (med_is_homog(m)) return bbox_inside(m.p0, m.p1, po) ? 1.f : 0.f;
    if (nsamples) (*nsamples)++;
    return grid_density(m, po);
}
// PhaseHG core/volume.cpp:150-154 (powf(x, 1.5f) evaluated as x*sqrt(x): <= 1 ulp apart; exact for g == 0)
__device__ __forceinline__ float phase_hg(v3 w, v3 wp, float g) {
    float costheta = vdot(w, wp);
    float x = 1.f + g * g - 2.f * g * costheta;
    return __fdiv_rn(1.f / (4.f * PV_PI_F) * (1.f - g * g), x * __fsqrt_rn(x));
}
template <class Med>
__device__ __forceinline__ float med_phase(const Med &m, v3 p, v3 w, v3 wp) {     // homogeneous.h:74-77, volume.h:92-94
    if (med_is_homog(m) && !bbox_inside(m.p0, m.p1, med_to_volume_p(m, p))) return 0.f;
    return phase_hg(w, wp, m.g);
}
// tau as a scalar s such that tau[b] = (sigma_a[b] + sigma_s[b]) * s.
// homogeneous.h:78-82: s = Distance(ray(t0), ray(t1)); core/volume.cpp:296-310: s = (sum of densities) * stepSize.
// (The reference accumulates sigma_t*density per bin before multiplying by stepSize; factoring the constant
// spectrum out changes rounding by O(1e-7) relative -- inside the 1e-4 per-ray tolerance, see DESIGN.md.)
template <class Med>
__device__ __forceinline__ float med_tau_scalar(const Med &m, v3 o, v3 d, float mint, float maxt, float stepSize,
                                                float u, uint32_t *nsamples) {
    float t0, t1;
    if (med_is_homog(m)) {
        if (!med_intersectp(m, o, d, mint, maxt, &t0, &t1)) return 0.f;
        return vlen(ray_at(o, d, t0) - ray_at(o, d, t1));
    }
    float length = vlen(d);
    if (length == 0.f) return 0.f;
    v3 dn = vdiv(d, length);
    if (!med_intersectp(m, o, dn, mint * length, maxt * length, &t0, &t1)) return 0.f;
    float s = 0.f;
    t0 += u * stepSize;
    while (t0 < t1) {
        s += grid_density(m, med_to_volume_p(m, ray_at(o, dn, t0)));
        if (nsamples) (*nsamples)++;
        t0 += stepSize;
    }
    return s * stepSize;
}

// ---------------------------------------------------------------- BVH + triangles
struct NodeRaw { float4 a, b; };
__device__ __forceinline__ bool node_slab(const float *b, v3 o, float mint, float maxt, v3 invDir, const int *neg) {
    // accelerators/bvh.cpp:167-189
    float tmin = (b[neg[0] ? 3 : 0] - o.x) * invDir.x;
    float tmax = (b[neg[0] ? 0 : 3] - o.x) * invDir.x;
    float tymin = (b[neg[1] ? 4 : 1] - o.y) * invDir.y;
    float tymax = (b[neg[1] ? 1 : 4] - o.y) * invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (b[neg[2] ? 5 : 2] - o.z) * invDir.z;
    float tzmax = (b[neg[2] ? 2 : 5] - o.z) * invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return (tmin < maxt) && (tmax > mint);
}
// shapes/trianglemesh.cpp:127-158 / :211-241
__device__ __forceinline__ bool tri_hit(const float *tv, v3 o, v3 d, float mint, float maxt, float *tHit) {
    v3 p1 = V3(__ldg(tv + 0), __ldg(tv + 1), __ldg(tv + 2));
    v3 p2 = V3(__ldg(tv + 3), __ldg(tv + 4), __ldg(tv + 5));
    v3 p3 = V3(__ldg(tv + 6), __ldg(tv + 7), __ldg(tv + 8));
    v3 e1 = p2 - p1, e2 = p3 - p1;
    v3 s1 = vcross(d, e2);
    float divisor = vdot(s1, e1);
    if (divisor == 0.f) return false;
    float invDivisor = __fdiv_rn(1.f, divisor);
    v3 s = o - p1;
    float b1 = vdot(s, s1) * invDivisor;
    if (b1 < 0.f || b1 > 1.f) return false;
    v3 s2 = vcross(s, e1);
    float b2 = vdot(d, s2) * invDivisor;
    if (b2 < 0.f || b1 + b2 > 1.f) return false;
    float t = vdot(e2, s2) * invDivisor;
    if (t < mint || t > maxt) return false;
    *tHit = t;
    return true;
}
// shapes/sphere.cpp:58-110 (Intersect) == :167-214 (IntersectP): the hit parameter.  Quadratic: core/pbrt.h:309-323.
__device__ __forceinline__ void sphere_phit(const pv_sphere &s, v3 ro, v3 rd, float thit, v3 *phit, float *phi) {
    v3 p = ray_at(ro, rd, thit);
    if (p.x == 0.f && p.y == 0.f) p.x = 1e-5f * s.radius;
    float ph = atan2f(p.y, p.x);
    if (ph < 0.f) ph = (float)((double)ph + 2.0 * 3.14159265358979323846);
    *phit = p; *phi = ph;
}
__device__ __forceinline__ bool sphere_clipped(const pv_sphere &s, v3 phit, float phi) {
    return (s.zmin > -s.radius && phit.z < s.zmin) || (s.zmax < s.radius && phit.z > s.zmax) || phi > s.phi_max;
}
static __device__ __noinline__ bool sphere_hit(const pv_sphere *sp, float ox, float oy, float oz, float dx, float dy, float dz, float mint, float maxt,
                                        float *tHit) {
    const pv_sphere &s = *sp;
    const v3 ro = xf_point(s.world_to_object, V3(ox, oy, oz)), rd = xf_vec(s.world_to_object, V3(dx, dy, dz));
    const float A = rd.x * rd.x + rd.y * rd.y + rd.z * rd.z;
    const float B = 2.f * (rd.x * ro.x + rd.y * ro.y + rd.z * ro.z);
    const float C = ro.x * ro.x + ro.y * ro.y + ro.z * ro.z - s.radius * s.radius;
    const float discrim = B * B - 4.f * A * C;
    if (discrim < 0.f) return false;
    const float rootDiscrim = __fsqrt_rn(discrim);
    const float q = B < 0.f ? -.5f * (B - rootDiscrim) : -.5f * (B + rootDiscrim);
    float t0 = __fdiv_rn(q, A), t1 = __fdiv_rn(C, q);
    if (t0 > t1) { const float t = t0; t0 = t1; t1 = t; }
    if (t0 > maxt || t1 < mint) return false;
    float thit = t0;
    if (t0 < mint) { thit = t1; if (thit > maxt) return false; }
    v3 phit; float phi;
    sphere_phit(s, ro, rd, thit, &phit, &phi);
    if (sphere_clipped(s, phit, phi)) {
        if (thit == t1) return false;
        if (t1 > maxt) return false;
        thit = t1;
        sphere_phit(s, ro, rd, thit, &phit, &phi);
        if (sphere_clipped(s, phit, phi)) return false;
    }
    *tHit = thit;
    return true;
}
// shapes/sphere.cpp:112-163 + core/diffgeom.cpp:40-55: hit point, normal and dpdu (world space), rayEpsilon
static __device__ __noinline__ void sphere_dg(const pv_sphere *sp, v3 o, v3 d, float t, v3 *hp, v3 *nn, v3 *dpdu_w, float *eps) {
    const pv_sphere &s = *sp;
    const v3 ro = xf_point(s.world_to_object, o), rd = xf_vec(s.world_to_object, d);
    v3 phit; float phi;
    sphere_phit(s, ro, rd, t, &phit, &phi);
    const float theta = acosf(fminf(fmaxf(__fdiv_rn(phit.z, s.radius), -1.f), 1.f));
    const float zradius = __fsqrt_rn(phit.x * phit.x + phit.y * phit.y);
    const float invzradius = __fdiv_rn(1.f, zradius);
    const float cosphi = phit.x * invzradius, sinphi = phit.y * invzradius;
    const v3 dpdu = V3(-s.phi_max * phit.y, s.phi_max * phit.x, 0.f);
    const v3 dpdv = V3(phit.z * cosphi, phit.z * sinphi, -s.radius * sinf(theta)) * (s.theta_max - s.theta_min);
    const v3 wu = xf_vec(s.object_to_world, dpdu), wv = xf_vec(s.object_to_world, dpdv);
    v3 n = vnorm(vcross(wu, wv));
    if (s.flip_normal) n = n * -1.f;
    *hp = xf_point(s.object_to_world, phit); *nn = n; *dpdu_w = wu; *eps = 5e-4f * t;
}
struct BvhCounters { uint32_t nodes, tris; };
// accelerators/bvh.cpp:585-636 (ANY = false) and :639-685 (ANY = true)
// SPH: the scene holds sphere primitives (the NaN tag is only looked for then; kernels are instantiated for both cases so
// that triangle-only scenes run exactly the triangle-only code)
template <bool ANY, bool SPH>
__device__ __forceinline__ int bvh_traverse(const DevScene &sc, v3 o, v3 d, float mint, float *maxt, BvhCounters *bc) {
    if (!sc.n_nodes) return -1;
    int hit = -1;
    v3 invDir = V3(__fdiv_rn(1.f, d.x), __fdiv_rn(1.f, d.y), __fdiv_rn(1.f, d.z));
    int neg[3] = {invDir.x < 0, invDir.y < 0, invDir.z < 0};
    uint32_t todoOffset = 0, nodeNum = 0, todo[64];
    for (;;) {
        const float4 *np = reinterpret_cast<const float4 *>(sc.nodes + nodeNum);
        float4 a = __ldg(np), b = __ldg(np + 1);
        float bounds[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
        uint32_t offset = __float_as_uint(b.z);
        uint32_t meta = __float_as_uint(b.w);
        uint32_t nPrims = meta & 0xffu, axis = (meta >> 8) & 0xffu;
        if (bc) bc->nodes++;
        if (node_slab(bounds, o, mint, *maxt, invDir, neg)) {
            if (nPrims > 0) {
                for (uint32_t i = 0; i < nPrims; ++i) {
                    float t;
                    if (bc) bc->tris++;
                    const float *tv = sc.tri + 9 * (size_t)(offset + i);
                    bool ph;
                    if (SPH) {
                        const float tag = __ldg(tv);
                        ph = tag == tag ? tri_hit(tv, o, d, mint, *maxt, &t)
                                        : sphere_hit(sc.spheres + (__float_as_uint(tag) & PV_SPHERE_INDEX_MASK), o.x, o.y, o.z, d.x, d.y, d.z, mint, *maxt, &t);
                    } else ph = tri_hit(tv, o, d, mint, *maxt, &t);
                    if (ph) {
                        if (ANY) return (int)(offset + i);
                        hit = (int)(offset + i);
                        *maxt = t;
                    }
                }
                if (todoOffset == 0) break;
                nodeNum = todo[--todoOffset];
            } else {
                if (neg[axis]) { todo[todoOffset++] = nodeNum + 1; nodeNum = offset; }
                else { todo[todoOffset++] = offset; nodeNum = nodeNum + 1; }
            }
        } else {
            if (todoOffset == 0) break;
            nodeNum = todo[--todoOffset];
        }
    }
    return hit;
}

// ---------------------------------------------------------------- lights (point query)
__device__ __forceinline__ float spot_falloff(const pv_light &l, v3 w) {                  // lights/spot.cpp:60-69
    v3 wl = vnorm(xf_vec(l.world_to_light, w));
    float costheta = wl.z;
    if (costheta < l.cos_total_width) return 0.f;
    if (costheta > l.cos_falloff_start) return 1.f;
    float delta = __fdiv_rn(costheta - l.cos_total_width, l.cos_falloff_start - l.cos_total_width);
    return delta * delta * delta * delta;
}
struct LightQuery {          // Sample_L(p, ...) reduced to scalars: L[b] = intensity[b] * scale  (or (I*falloff)/d2)
    v3 wi; float falloff; float inv_mode_d2; int point_like; v3 vis_o, vis_d; float vis_mint, vis_maxt;
};
// lights/point.cpp:50-57, spot.cpp:50-57, distant.cpp:48-55, VisibilityTester core/light.h:85-101
__device__ __forceinline__ void light_query(const pv_light &l, v3 p, LightQuery *q) {
    q->vis_o = p; q->vis_mint = 0.f;
    if (l.type == PV_LIGHT_DISTANT) {
        q->wi = V3(l.dir[0], l.dir[1], l.dir[2]);
        q->vis_d = q->wi; q->vis_maxt = INFINITY; q->falloff = 1.f; q->inv_mode_d2 = 1.f; q->point_like = 0;
        return;
    }
    v3 lp = V3(l.pos[0], l.pos[1], l.pos[2]);
    q->wi = vnorm(lp - p);
    float dist = vlen(p - lp);
    q->vis_d = vdiv(lp - p, dist); q->vis_maxt = dist * (1.f - 0.f);
    q->inv_mode_d2 = dist2(lp, p);
    q->point_like = 1;
    q->falloff = (l.type == PV_LIGHT_SPOT) ? spot_falloff(l, -q->wi) : 1.f;
}
// radiance for bin value I: point -> I / d2; spot -> (I * falloff) / d2; distant -> I
__device__ __forceinline__ float light_L_bin(const pv_light &l, const LightQuery &q, float I) {
    if (!q.point_like) return I;
    if (l.type == PV_LIGHT_SPOT) return __fdiv_rn(I * q.falloff, q.inv_mode_d2);
    return __fdiv_rn(I, q.inv_mode_d2);
}

// ---------------------------------------------------------------- DiffuseAreaLight over triangles
// Triangle::Intersect's DifferentialGeometry normal for default uvs (shapes/trianglemesh.cpp:160-205, core/diffgeom.cpp:40-55)
__device__ __forceinline__ v3 tri_dg_nn(const float *tv, uint32_t flags) {
    const v3 p1 = V3(tv[0], tv[1], tv[2]), p2 = V3(tv[3], tv[4], tv[5]), p3 = V3(tv[6], tv[7], tv[8]);
    const v3 dp1 = p1 - p3, dp2 = p2 - p3;
    const v3 dpdu = (dp1 * -1.f - dp2 * -1.f) * 1.f;                       // (dv2*dp1 - dv1*dp2) * invdet, dv1 = dv2 = -1, det = 1
    const v3 dpdv = (dp1 * -0.f + dp2 * -1.f) * 1.f;                       // (-du2*dp1 + du1*dp2) * invdet, du2 = 0, du1 = -1
    v3 nn = vnorm(vcross(dpdu, dpdv));
    if (((flags & PV_AREA_REVERSE_ORIENTATION) != 0) ^ ((flags & PV_AREA_SWAPS_HANDEDNESS) != 0)) nn = nn * -1.f;
    return nn;
}
// the triangle ShapeSet::Sample picks for the component sample u (SampleDiscrete: upper_bound on the CDF - 1), its point for
// (u0, u1) (Triangle::Sample shapes/trianglemesh.cpp:444-456) and the geometric normal there
__device__ __forceinline__ v3 area_sample_point(const DevScene &sc, const pv_light &l, int li, float uComp, float u0, float u1, v3 *ns) {
    const int n = (int)l.area.n_tris;
    const float *cdf = sc.ltri_cdf + sc.lcdf_off[li];
    int lo = 0, hi = n + 1;
    while (lo < hi) { const int mid = (lo + hi) / 2; if (uComp < cdf[mid]) hi = mid; else lo = mid + 1; }
    const int sn = max(lo - 1, 0);
    const float *tv = sc.ltris + 9 * (size_t)(l.area.first_tri + sn);
    const v3 p1 = V3(tv[0], tv[1], tv[2]), p2 = V3(tv[3], tv[4], tv[5]), p3 = V3(tv[6], tv[7], tv[8]);
    const float su1 = __fsqrt_rn(u0), b1 = 1.f - su1, b2 = u1 * su1;
    *ns = vnorm(vcross(p2 - p1, p3 - p1));
    if (l.area.flags & PV_AREA_REVERSE_ORIENTATION) *ns = *ns * -1.f;
    return (p1 * b1 + p2 * b2) + p3 * (1.f - b1 - b2);
}
// DiffuseAreaLight::Sample_L(p, pEpsilon, ls, time, &wi, &pdf, &vis) (lights/diffuse.cpp:69-86) with ShapeSet::Sample(p, ls, Ns)
// (core/light.cpp:139-158: ONE ray towards the sampled point against every shape, hit and normal of the LAST shape it hits) and
// ShapeSet::Pdf(p, wi) (:167-172 over Shape::Pdf core/shape.cpp:86-99).  Returns whether the light faces p (L = Lemit, else 0).
static __device__ __noinline__ bool area_sample_L(const DevScene &sc, int li, v3 p, float uComp, float u0, float u1, v3 *wi, float *pdf,
                                           v3 *vis_d, float *vis_maxt) {
    const pv_light &l = sc.lights[li];
    const uint32_t n = l.area.n_tris;
    *pdf = 0.f; *wi = V3(0.f, 0.f, 1.f); *vis_d = *wi; *vis_maxt = 0.f;
    if (!n) return false;
    const float *tris = sc.ltris + 9 * (size_t)l.area.first_tri, *area = sc.ltri_area + l.area.first_tri;
    v3 ns;
    const v3 pt = area_sample_point(sc, l, li, uComp, u0, u1, &ns);
    const v3 rdir = pt - p;
    float thit = 1.f;
    for (uint32_t i = 0; i < n; ++i) {
        float t;
        if (tri_hit(tris + 9 * (size_t)i, p, rdir, 1e-3f, INFINITY, &t)) { thit = t; ns = tri_dg_nn(tris + 9 * (size_t)i, l.area.flags); }
    }
    const v3 ps = ray_at(p, rdir, thit);
    *wi = vnorm(ps - p);
    float pdfsum = 0.f;
    for (uint32_t i = 0; i < n; ++i) {
        float t;
        if (!tri_hit(tris + 9 * (size_t)i, p, *wi, 1e-3f, INFINITY, &t)) continue;
        const v3 nn = tri_dg_nn(tris + 9 * (size_t)i, l.area.flags);
        float pd = __fdiv_rn(dist2(p, ray_at(p, *wi, t)), fabsf(vdot(nn, -(*wi))) * area[i]);
        if (isinf(pd)) pd = 0.f;
        pdfsum += area[i] * pd;
    }
    *pdf = __fdiv_rn(pdfsum, sc.larea_sum[li]);
    // visibility->SetSegment(p, pEpsilon = 0, ps, 1e-3f, time) (core/light.h:85-93)
    const float dist = vlen(p - ps);
    *vis_d = vdiv(ps - p, dist); *vis_maxt = dist * (1.f - 1e-3f);
    return vdot(ns, -(*wi)) > 0.f;                                       // DiffuseAreaLight::L (lights/diffuse.h:51-53)
}

// core/spectrum.h:433-439 for a per-thread 30-float array
__device__ __forceinline__ float spec_y(const float *cie_y, const float *c) {
    float yy = 0.f;
#pragma unroll
    for (int i = 0; i < PV_NSPEC; ++i) yy += cie_y[i] * c[i];
    return __fdiv_rn(yy * 300.f, 106.856895f * (float)PV_NSPEC);
}

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
