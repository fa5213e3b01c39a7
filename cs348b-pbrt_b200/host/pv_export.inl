// pv_export.inl -- flatten the reference's host scene objects into the C-ABI scene description
// (include/pv.h pv_scene_desc), SURVEY.md Appendix B.
//
// Included by a translation unit that has ALREADY included the reference headers with
// `#define private public` / `#define protected public` in effect (core/scene.h, accelerators/bvh.h,
// shapes/trianglemesh.h, lights/{point,spot,distant}.h, volumes/{homogeneous,volumegrid,rainbow}.h,
// materials/{matte,glass}.h) and include/pv.h.  Users: host/pv_pbrt_adapter.cpp (the drop-in classes) and
// oracle/ref_harness.cpp (fixture generation).  No reference source is modified.
#pragma once
#include <map>
#include <string>
#include <vector>
#include <stdio.h>
#include <string.h>

// Layout of accelerators/bvh.cpp:154-164 (the struct is private to that .cpp, so it is re-declared here).
struct LinearBVHNode {
    BBox bounds;
    union { uint32_t primitivesOffset; uint32_t secondChildOffset; };
    uint8_t nPrimitives, axis, pad[2];
};

struct PvHostScene {
    std::vector<pv_bvh_node> nodes;
    std::vector<float> tri;
    std::vector<uint32_t> prim_material;
    std::vector<pv_material> materials;
    std::vector<pv_light> lights;
    std::vector<float> density;
    std::vector<uint32_t> prim_shape;          // PV_SHAPE_TRIANGLE or an index into spheres
    std::vector<pv_sphere> spheres;
    std::vector<float> light_tris;             // triangles of the DiffuseAreaLights' ShapeSets (PV_LIGHT_AREA), 9 floats each
    pv_medium medium;
    bool has_medium;
    // empty when the device description holds the scene's materials exactly; else what was approximated (the all-maps photon
    // pass refuses such a scene).  has_unknown_specular: a material other than matte / glass was met -- it may have specular
    // lobes, which matter even to the volume-only pass (a diffuse bounce ends a volume-only path whatever the BRDF is, Q6).
    std::string inexact;
    bool has_unknown_specular;
    pv_scene_desc desc;
};

static inline void pv_spec_out(const Spectrum &s, float *dst) { memcpy(dst, s.c, sizeof(float) * nSpectralSamples); }
static inline void pv_mat_out(const Transform &t, float *dst) {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) dst[4 * i + j] = t.m.m[i][j];
}

// medium_only: export the volume region and nothing else (no geometry, no lights) -- all that EmissionIntegrator::Li reads
// (integrators/emission.cpp:63-106); lets "emission", pbrt's DEFAULT volume integrator (core/api.cpp:211), run on the device in
// scenes whose surfaces or lights are off this path (area lights, other shapes, other accelerators).
// keep_unknown_lights (the golden-vector harness only, never the drop-in): a light the device path does not know keeps its
// slot in the light list as a placeholder of type PV_LIGHT_UNKNOWN_SLOT (power filled in) instead of failing the export, so the
// oracle can be given such lights on the side under the reference's own light indices (area lights).
#define PV_LIGHT_UNKNOWN_SLOT 100
// A BVH builder the exporter can call for a scene whose aggregate does not hold a LinearBVHNode array (Accelerator "kdtree" /
// "grid"), or when the caller wants the device's tree in place of the reference's (PV_BVH=gpu): the adapter passes one that calls
// pv_build_bvh on its context.  bounds: 6 floats per primitive; fills nodes and order (position in the reordered primitive array
// -> index into the list the bounds came from).
struct PvBvhBuilder {
    virtual bool build(const float *bounds, uint32_t n, uint32_t max_prims_in_node, std::vector<pv_bvh_node> &nodes,
                       std::vector<uint32_t> &order, std::string &err) = 0;
    bool force;                 // use it even when the aggregate is a BVHAccel
    PvBvhBuilder() : force(false) {}
    virtual ~PvBvhBuilder() {}
};

static bool pv_export_scene(const Scene *scene, PvHostScene &hs, std::string &err, bool medium_only = false,
                            bool keep_unknown_lights = false, PvBvhBuilder *builder = NULL) {
    BVHAccel *bvh = medium_only ? NULL : dynamic_cast<BVHAccel *>(scene->aggregate);
    // the fully refined primitives and, per position of the device's primitive array, which of them sits there
    std::vector<Reference<Primitive> > refined;
    const std::vector<Reference<Primitive> > *prims = bvh ? &bvh->primitives : NULL;
    std::vector<uint32_t> order;
    hs.nodes.clear();
    if (!medium_only && (!bvh || (builder && builder->force))) {
        if (!builder) { err = "pv: the scene aggregate is not the \"bvh\" accelerator (and no device BVH builder was given)"; return false; }
        uint32_t max_prims = 4;                                   // the reference's "maxnodeprims" default (accelerators/bvh.cpp:694)
        if (bvh) max_prims = bvh->maxPrimsInNode;
        else if (KdTreeAccel *kd = dynamic_cast<KdTreeAccel *>(scene->aggregate)) prims = &kd->primitives;   // refined by its ctor (kdtreeaccel.cpp:96-98)
        else if (GridAccel *ga = dynamic_cast<GridAccel *>(scene->aggregate)) {
            for (size_t i = 0; i < ga->primitives.size(); ++i) ga->primitives[i]->FullyRefine(refined);       // grid.cpp:46-51 may leave them unrefined
            prims = &refined;
        } else { err = "pv: the scene aggregate is none of the \"bvh\", \"kdtree\", \"grid\" accelerators"; return false; }
        std::vector<float> bounds(6 * prims->size());
        for (size_t i = 0; i < prims->size(); ++i) {
            BBox b = (*prims)[i]->WorldBound();
            bounds[6 * i + 0] = b.pMin.x; bounds[6 * i + 1] = b.pMin.y; bounds[6 * i + 2] = b.pMin.z;
            bounds[6 * i + 3] = b.pMax.x; bounds[6 * i + 4] = b.pMax.y; bounds[6 * i + 5] = b.pMax.z;
        }
        if (!builder->build(bounds.data(), (uint32_t)prims->size(), max_prims, hs.nodes, order, err)) return false;
    } else if (bvh) {
        const LinearBVHNode *nodes = (const LinearBVHNode *)bvh->nodes;
        uint32_t nNodes = 0;
        if (nodes) {   // depth-first layout (flattenBVHTree bvh.cpp:559-577): the node count is the largest index reached + 1
            std::vector<uint32_t> todo; todo.push_back(0);
            while (!todo.empty()) {
                uint32_t i = todo.back(); todo.pop_back();
                if (i + 1 > nNodes) nNodes = i + 1;
                if (nodes[i].nPrimitives == 0) { todo.push_back(i + 1); todo.push_back(nodes[i].secondChildOffset); }
            }
        }
        hs.nodes.resize(nNodes);
        if (nNodes) memcpy(hs.nodes.data(), nodes, sizeof(pv_bvh_node) * nNodes);
    }
    const uint32_t nPrims = prims ? (uint32_t)prims->size() : 0u;
    const uint32_t nNodes = (uint32_t)hs.nodes.size();
    if (order.empty()) { order.resize(nPrims); for (uint32_t i = 0; i < nPrims; ++i) order[i] = i; }
    hs.tri.assign(9 * (size_t)nPrims, 0.f);
    hs.prim_material.assign(nPrims, 0u);
    hs.materials.clear();
    hs.prim_shape.assign(nPrims, PV_SHAPE_TRIANGLE);
    hs.spheres.clear();
    hs.inexact.clear(); hs.has_unknown_specular = false;
    std::map<const Material *, uint32_t> matIndex;
    DifferentialGeometry dummy;
    bool warned_uv = false;
    for (uint32_t i = 0; i < nPrims; ++i) {
        const GeometricPrimitive *gp = dynamic_cast<const GeometricPrimitive *>((*prims)[order[i]].GetPtr());
        if (!gp) { err = "pv: a primitive is not a GeometricPrimitive (instancing is out of scope)"; return false; }
        const Triangle *t = dynamic_cast<const Triangle *>(gp->shape.GetPtr());
        const Sphere *sph = t ? NULL : dynamic_cast<const Sphere *>(gp->shape.GetPtr());
        if (!t && !sph) { err = "pv: a shape is neither a Triangle nor a Sphere (triangle meshes and spheres are on this path)"; return false; }
        if (sph) {                                          // shapes/sphere.cpp:40-50
            pv_sphere ps; memset(&ps, 0, sizeof(ps));
            pv_mat_out(*sph->ObjectToWorld, ps.object_to_world); pv_mat_out(*sph->WorldToObject, ps.world_to_object);
            ps.radius = sph->radius; ps.zmin = sph->zmin; ps.zmax = sph->zmax;
            ps.theta_min = sph->thetaMin; ps.theta_max = sph->thetaMax; ps.phi_max = sph->phiMax;
            ps.flip_normal = (sph->ReverseOrientation ^ sph->TransformSwapsHandedness) ? 1 : 0;
            hs.prim_shape[i] = (uint32_t)hs.spheres.size(); hs.spheres.push_back(ps);
        }
        if (t && t->mesh->uvs && !warned_uv) {
            warned_uv = true;
            fprintf(stderr, "pv: warning: explicit triangle uvs are ignored by the photon shooter (default-uv shading frame, shapes/trianglemesh.cpp:166-175); "
                            "diffuse bounce directions are distributed identically but drawn in another tangent frame\n");
        }
        for (int k = 0; t && k < 3; ++k) {
            const Point &p = t->mesh->p[t->v[k]];          // already world space (shapes/trianglemesh.cpp:70-71)
            hs.tri[9 * i + 3 * k + 0] = p.x; hs.tri[9 * i + 3 * k + 1] = p.y; hs.tri[9 * i + 3 * k + 2] = p.z;
        }
        const Material *m = gp->material.GetPtr();
        if (!matIndex.count(m)) {
            pv_material pm; memset(&pm, 0, sizeof(pm));
            if (const MatteMaterial *mm = dynamic_cast<const MatteMaterial *>(m)) {
                pm.type = PV_MAT_MATTE;
                pv_spec_out(mm->Kd->Evaluate(dummy).Clamp(), pm.kd);
                // the device material is ONE reflectance: a spatially varying Kd texture is flattened to its value at (u, v) = (0, 0)
                DifferentialGeometry probe; probe.u = 0.37f; probe.v = 0.61f; probe.p = Point(0.37f, 0.61f, 0.13f);
                if (mm->Kd->Evaluate(probe).Clamp() != mm->Kd->Evaluate(dummy).Clamp()) {
                    fprintf(stderr, "pv: warning: textured matte Kd on primitive %u is flattened to its value at uv = (0, 0) for photon shooting\n", i);
                    if (hs.inexact.empty()) hs.inexact = "a matte material has a textured Kd";
                }
                if (mm->sigma->Evaluate(dummy) != 0.f) {
                    fprintf(stderr, "pv: warning: matte sigma != 0 (Oren-Nayar) is treated as Lambertian\n");
                    if (hs.inexact.empty()) hs.inexact = "a matte material has sigma != 0 (Oren-Nayar)";
                }
            } else if (const GlassMaterial *gm = dynamic_cast<const GlassMaterial *>(m)) {
                pm.type = PV_MAT_GLASS;
                pv_spec_out(gm->Kr->Evaluate(dummy).Clamp(), pm.kr);
                pv_spec_out(gm->Kt->Evaluate(dummy).Clamp(), pm.kt);
                pm.index = gm->index->Evaluate(dummy);
                pm.vn = gm->Vn;
            } else {
                fprintf(stderr, "pv: warning: unsupported material on primitive %u, photons treat it as black matte\n", i);
                pm.type = PV_MAT_MATTE;
                hs.inexact = "a material is neither matte nor glass"; hs.has_unknown_specular = true;
            }
            matIndex[m] = (uint32_t)hs.materials.size(); hs.materials.push_back(pm);
        }
        hs.prim_material[i] = matIndex[m];
    }
    hs.lights.clear(); hs.light_tris.clear();
    for (size_t i = 0; !medium_only && i < scene->lights.size(); ++i) {
        pv_light pl; memset(&pl, 0, sizeof(pl));
        Light *l = scene->lights[i];
        pv_mat_out(l->LightToWorld, pl.light_to_world);
        pv_mat_out(l->WorldToLight, pl.world_to_light);
        pl.power_y = l->Power(scene).y();
        if (PointLight *p = dynamic_cast<PointLight *>(l)) {
            pl.type = PV_LIGHT_POINT;
            pl.pos[0] = p->lightPos.x; pl.pos[1] = p->lightPos.y; pl.pos[2] = p->lightPos.z;
            pv_spec_out(p->Intensity, pl.intensity);
        } else if (SpotLight *s = dynamic_cast<SpotLight *>(l)) {
            pl.type = PV_LIGHT_SPOT;
            pl.pos[0] = s->lightPos.x; pl.pos[1] = s->lightPos.y; pl.pos[2] = s->lightPos.z;
            pv_spec_out(s->Intensity, pl.intensity);
            pl.cos_total_width = s->cosTotalWidth; pl.cos_falloff_start = s->cosFalloffStart;
        } else if (DistantLight *d = dynamic_cast<DistantLight *>(l)) {
            pl.type = PV_LIGHT_DISTANT;
            pl.dir[0] = d->lightDir.x; pl.dir[1] = d->lightDir.y; pl.dir[2] = d->lightDir.z;
            pv_spec_out(d->L, pl.intensity);
        } else if (DiffuseAreaLight *al = keep_unknown_lights ? NULL : dynamic_cast<DiffuseAreaLight *>(l)) {
            // lights/diffuse.cpp:39-106: Lemit and the triangles of its ShapeSet in the ShapeSet's own (refine) order -- the order the
            // area distribution samples by (core/light.cpp:114-137)
            pl.type = PV_LIGHT_AREA;
            pv_spec_out(al->Lemit, pl.intensity);
            pl.area.first_tri = (uint32_t)(hs.light_tris.size() / 9); pl.area.n_tris = (uint32_t)al->shapeSet->shapes.size(); pl.area.flags = 0;
            for (uint32_t k = 0; k < pl.area.n_tris; ++k) {
                const Triangle *t = dynamic_cast<const Triangle *>(al->shapeSet->shapes[k].GetPtr());
                if (!t) { err = "pv: an area light over a shape that is not a triangle mesh (triangle-mesh area lights are on this path)"; return false; }
                pl.area.flags = (t->ReverseOrientation ? PV_AREA_REVERSE_ORIENTATION : 0u) | (t->TransformSwapsHandedness ? PV_AREA_SWAPS_HANDEDNESS : 0u);
                for (int c = 0; c < 3; ++c) { const Point &q = t->mesh->p[t->v[c]]; hs.light_tris.push_back(q.x); hs.light_tris.push_back(q.y); hs.light_tris.push_back(q.z); }
            }
            if (dynamic_cast<RainbowVolume *>(scene->volumeRegion)) { err = "pv: an area light in a rainbow medium is not on this path"; return false; }
        } else if (keep_unknown_lights) pl.type = PV_LIGHT_UNKNOWN_SLOT;
        else { err = "pv: unsupported light type (point, spot, distant and triangle-mesh area lights are on this path)"; return false; }
        hs.lights.push_back(pl);
    }
    hs.has_medium = false;
    memset(&hs.medium, 0, sizeof(hs.medium));
    hs.density.clear();
    if (VolumeRegion *vr = scene->volumeRegion) {
        pv_medium &m = hs.medium;
        // RainbowVolume derives from HomogeneousVolumeDensity: test it first
        if (HomogeneousVolumeDensity *h = dynamic_cast<HomogeneousVolumeDensity *>(vr)) {
            m.type = dynamic_cast<RainbowVolume *>(vr) ? PV_MEDIUM_RAINBOW : PV_MEDIUM_HOMOGENEOUS;
            pv_mat_out(h->WorldToVolume, m.world_to_volume);
            m.p0[0] = h->extent.pMin.x; m.p0[1] = h->extent.pMin.y; m.p0[2] = h->extent.pMin.z;
            m.p1[0] = h->extent.pMax.x; m.p1[1] = h->extent.pMax.y; m.p1[2] = h->extent.pMax.z;
            pv_spec_out(h->sig_a, m.sigma_a); pv_spec_out(h->sig_s, m.sigma_s); pv_spec_out(h->le, m.le); m.g = h->g;
        } else if (VolumeGridDensity *gd = dynamic_cast<VolumeGridDensity *>(vr)) {
            m.type = PV_MEDIUM_GRID;
            pv_mat_out(gd->WorldToVolume, m.world_to_volume);
            m.p0[0] = gd->extent.pMin.x; m.p0[1] = gd->extent.pMin.y; m.p0[2] = gd->extent.pMin.z;
            m.p1[0] = gd->extent.pMax.x; m.p1[1] = gd->extent.pMax.y; m.p1[2] = gd->extent.pMax.z;
            pv_spec_out(gd->sig_a, m.sigma_a); pv_spec_out(gd->sig_s, m.sigma_s); pv_spec_out(gd->le, m.le); m.g = gd->g;
            m.nx = gd->nx; m.ny = gd->ny; m.nz = gd->nz;
            hs.density.assign(gd->density, gd->density + (size_t)gd->nx * gd->ny * gd->nz);
            m.density = hs.density.data();
        } else if (ExponentialDensity *ed = dynamic_cast<ExponentialDensity *>(vr)) {
            m.type = PV_MEDIUM_EXPONENTIAL;                     // volumes/exponential.h:43-68: {a, b, updir} travel in the density slot
            pv_mat_out(ed->WorldToVolume, m.world_to_volume);
            m.p0[0] = ed->extent.pMin.x; m.p0[1] = ed->extent.pMin.y; m.p0[2] = ed->extent.pMin.z;
            m.p1[0] = ed->extent.pMax.x; m.p1[1] = ed->extent.pMax.y; m.p1[2] = ed->extent.pMax.z;
            pv_spec_out(ed->sig_a, m.sigma_a); pv_spec_out(ed->sig_s, m.sigma_s); pv_spec_out(ed->le, m.le); m.g = ed->g;
            m.nx = 5; m.ny = 1; m.nz = 1;
            const float prm[5] = {ed->a, ed->b, ed->upDir.x, ed->upDir.y, ed->upDir.z};
            hs.density.assign(prm, prm + 5);
            m.density = hs.density.data();
        } else { err = "pv: unsupported volume region (homogeneous, volumegrid, exponential and rainbow are on this path)"; return false; }
        hs.has_medium = true;
    }
    pv_scene_desc &d = hs.desc;
    memset(&d, 0, sizeof(d));
    d.nodes = hs.nodes.data(); d.n_nodes = nNodes;
    d.tri_verts = hs.tri.data(); d.prim_material = hs.prim_material.data(); d.n_prims = nPrims;
    d.materials = hs.materials.data(); d.n_materials = (uint32_t)hs.materials.size();
    d.lights = hs.lights.data(); d.n_lights = (uint32_t)hs.lights.size();
    d.medium = hs.has_medium ? &hs.medium : NULL;
    const BBox &wb = scene->WorldBound();
    d.world_bound[0] = wb.pMin.x; d.world_bound[1] = wb.pMin.y; d.world_bound[2] = wb.pMin.z;
    d.world_bound[3] = wb.pMax.x; d.world_bound[4] = wb.pMax.y; d.world_bound[5] = wb.pMax.z;
    memcpy(d.cie_y, SampledSpectrum::Y.c, sizeof(d.cie_y));
    if (!hs.spheres.empty()) { d.prim_shape = hs.prim_shape.data(); d.spheres = hs.spheres.data(); d.n_spheres = (uint32_t)hs.spheres.size(); }
    if (!hs.light_tris.empty()) { d.light_tris = hs.light_tris.data(); d.n_light_tris = (uint32_t)(hs.light_tris.size() / 9); }
    return true;
}

// PVSCN001 interchange file (read by cs348b-pbrt_b200/sceneio.py)
static bool pv_write_scene_file(const PvHostScene &hs, const std::string &fn) {
    FILE *f = fopen(fn.c_str(), "wb");
    if (!f) { perror(fn.c_str()); return false; }
    const pv_scene_desc &d = hs.desc;
    uint32_t hdr[8] = {d.n_nodes, d.n_prims, d.n_materials, d.n_lights, hs.has_medium ? 1u : 0u, d.n_spheres, d.n_light_tris, 0};
    bool ok = fwrite("PVSCN001", 1, 8, f) == 8 && fwrite(hdr, 4, 8, f) == 8 && fwrite(d.world_bound, 4, 6, f) == 6 &&
              fwrite(d.cie_y, 4, PV_NSPEC, f) == PV_NSPEC;
    ok = ok && fwrite(hs.nodes.data(), sizeof(pv_bvh_node), hs.nodes.size(), f) == hs.nodes.size();
    ok = ok && fwrite(hs.tri.data(), 4, hs.tri.size(), f) == hs.tri.size();
    ok = ok && fwrite(hs.prim_material.data(), 4, hs.prim_material.size(), f) == hs.prim_material.size();
    ok = ok && fwrite(hs.materials.data(), sizeof(pv_material), hs.materials.size(), f) == hs.materials.size();
    ok = ok && fwrite(hs.lights.data(), sizeof(pv_light), hs.lights.size(), f) == hs.lights.size();
    if (ok && hs.has_medium) {
        const pv_medium &m = hs.medium;
        int32_t dims[3] = {m.nx, m.ny, m.nz};
        ok = fwrite(&m.type, 4, 1, f) == 1 && fwrite(m.world_to_volume, 4, 16, f) == 16 && fwrite(m.p0, 4, 3, f) == 3 &&
             fwrite(m.p1, 4, 3, f) == 3 && fwrite(m.sigma_a, 4, 30, f) == 30 && fwrite(m.sigma_s, 4, 30, f) == 30 &&
             fwrite(m.le, 4, 30, f) == 30 && fwrite(&m.g, 4, 1, f) == 1 && fwrite(dims, 4, 3, f) == 3;
        if (ok && (m.type == PV_MEDIUM_GRID || m.type == PV_MEDIUM_EXPONENTIAL)) ok = fwrite(hs.density.data(), 4, hs.density.size(), f) == hs.density.size();
    }
    if (ok && d.n_spheres)
        ok = fwrite(hs.prim_shape.data(), 4, hs.prim_shape.size(), f) == hs.prim_shape.size() &&
             fwrite(hs.spheres.data(), sizeof(pv_sphere), hs.spheres.size(), f) == hs.spheres.size();
    if (ok && d.n_light_tris) ok = fwrite(hs.light_tris.data(), 4, hs.light_tris.size(), f) == hs.light_tris.size();
    fclose(f);
    if (!ok) fprintf(stderr, "pv: short write on %s\n", fn.c_str());
    else fprintf(stderr, "[pv] exported scene: %u nodes, %u prims, %u materials, %u lights -> %s\n", d.n_nodes, d.n_prims, d.n_materials,
                 d.n_lights, fn.c_str());
    return ok;
}
