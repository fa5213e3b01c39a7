#!/usr/bin/env python
"""Generate tests/golden/*.npz and *.scn by running the REAL reference.

Needs /root/reference (this container only): builds oracle/_ref/ref_harness with
`make -C oracle ref`, runs it single-threaded (--ncores 1 => deterministic,
SURVEY.md 8c) and stores what it returns.  The GPU box never runs this script;
it only reads the committed outputs.

    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
from cs348b_pbrt_b200 import sceneio, scenes  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def run(scene_file, *ops):
    cmd = [HARNESS, scene_file] + [str(o) for o in ops]
    print("+", " ".join(cmd[:6]), "...")
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)


def query_points(rays, stepsize, n_per_ray, lo=-1.0, hi=1.0):
    """March points of camera rays inside the medium box (like Li's sample points)."""
    pts, ws = [], []
    lo = np.broadcast_to(np.asarray(lo, np.float64), (3,)); hi = np.broadcast_to(np.asarray(hi, np.float64), (3,))
    for r in rays:
        o, d = r["o"].astype(np.float64), r["d"].astype(np.float64)
        with np.errstate(divide="ignore"):
            t0 = ((lo - o) / d); t1 = ((hi - o) / d)
        tn = np.minimum(t0, t1).max(); tf = np.maximum(t0, t1).min()
        if tn >= tf:
            continue
        for k in range(n_per_ray):
            t = tn + (k + 0.5) * (tf - tn) / n_per_ray
            pts.append(o + d * t); ws.append(-d)
    return np.asarray(pts, np.float32), np.asarray(ws, np.float32)


def aggregate_test_rays(n, seed, bound=(-1.2, 1.2), target=None):
    """Ray set in the spirit of renderers/aggregatetest.cpp:61-119: random origins, sphere-uniform and
    axis-aligned directions, finite and infinite extents."""
    rng = np.random.default_rng(seed)
    o = rng.uniform(bound[0], bound[1], size=(n, 3)).astype(np.float32)
    z = rng.uniform(-1, 1, size=n); phi = rng.uniform(0, 2 * np.pi, size=n); r = np.sqrt(1 - z * z)
    d = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
    axis = rng.integers(0, 3, size=n); sign = rng.choice([-1.0, 1.0], size=n)
    ax = rng.random(n) < 0.15
    d[ax] = 0; d[ax, axis[ax]] = sign[ax]
    if target is not None:                                               # small object in a big scene: aim 40% of the rays at it
        aim = rng.random(n) < 0.4
        tgt = np.asarray(target[0], np.float32) + rng.uniform(-1, 1, size=(n, 3)).astype(np.float32) * np.asarray(target[1], np.float32)
        dd = tgt - o; dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        d[aim] = dd[aim]
    d *= rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)          # non-unit directions too
    rays = sceneio.make_rays(o, d, 0.0, np.inf)
    fin = rng.random(n) < 0.3
    rays["maxt"][fin] = rng.uniform(0.1, 2.0, size=fin.sum()).astype(np.float32)
    rays["mint"][rng.random(n) < 0.2] = np.float32(1e-3)
    # some rays start exactly on a wall
    on = rng.random(n) < 0.1
    rays["o"][on, 1] = np.float32(-1.0)
    return rays


def golden_for_scene(tmp, name, scene_file, nused, maxdist, stepsize, n_li_rays, li_res, knn_k_list, camera=None, box=(-1.0, 1.0),
                     wanted=0, shoot_step=0.05, hit_bound=(-1.2, 1.2), hit_target=None, q_near_photons=False, regions=False):
    out = {}
    camera = camera or {}
    scn = os.path.join(HERE, name + ".scn")
    pho = os.path.join(tmp, name + ".pho")
    stats = os.path.join(tmp, name + ".json")
    # regions: the scene has several Volume statements (AggregateVolume) -> one <name>.<i>.scn per region instead of <name>.scn
    export = ["--export-regions", os.path.join(HERE, name)] if regions else ["--export-scene", scn]
    run(scene_file, *export, "--shoot", "--dump-photons", pho, "--stats", stats)
    pos, wi, alpha = sceneio.read_photons(pho)
    st = json.load(open(stats))
    out["shot_pos"], out["shot_wi"], out["shot_alpha"] = pos, wi, alpha
    out["nshot"] = np.array([st["nshot"]], np.uint64)
    print("  %s: %d photons from %d paths" % (name, len(pos), st["nshot"]))

    # camera rays (+ jittered scatter sample) and query points along them
    rays = scenes.camera_rays(li_res, li_res, **camera)
    rng = np.random.default_rng(7)
    sel = rng.choice(len(rays), size=n_li_rays, replace=False)
    rays = rays[np.sort(sel)]
    rays["u_scatter"] = rng.random(len(rays)).astype(np.float32)
    rays["u_scatter"][:8] = np.float32(0.5)
    out["li_rays"] = rays
    pts, ws = query_points(rays[::4], stepsize, 6, box[0], box[1])
    if q_near_photons:          # photons fill a tiny part of the medium (a light beam): put the queries where the photons are
        pick = rng.choice(len(pos), size=len(pts), replace=False)
        pts = (pos[pick] + rng.uniform(-0.08, 0.08, size=(len(pts), 3))).astype(np.float32)
    out["q_pts"], out["q_w"] = pts, ws

    rf = os.path.join(tmp, "rays.bin"); qf = os.path.join(tmp, "q.bin")
    sceneio.write_rays(rf, rays); sceneio.write_queries(qf, pts, ws)
    ops = ["--load-photons", pho]
    for k, r2 in knn_k_list:
        ops += ["--knn", qf, k, repr(float(np.float32(r2))), os.path.join(tmp, "knn_%d.bin" % k)]
    ops += ["--lphoton", qf, os.path.join(tmp, "lph.bin"), "--li", rf, 1000, os.path.join(tmp, "li.bin"),
            "--transmittance", rf, 77, os.path.join(tmp, "tr.bin")]
    run(scene_file, *ops)
    for k, r2 in knn_k_list:
        nf, idx, d2 = sceneio.read_knn(os.path.join(tmp, "knn_%d.bin" % k))
        out["knn%d_nfound" % k], out["knn%d_idx" % k], out["knn%d_d2" % k] = nf, idx, d2
        out["knn%d_r2" % k] = np.array([r2], np.float32)
    out["lphoton_L"] = sceneio.read_spectra(os.path.join(tmp, "lph.bin"), b"PVSPEC01")[:, 0]
    li = sceneio.read_spectra(os.path.join(tmp, "li.bin"), b"PVLI0001", per=2)
    out["li_L"], out["li_T"] = li[:, 0], li[:, 1]
    with open(os.path.join(tmp, "tr.bin"), "rb") as f:
        buf = f.read()
    n = len(rays)
    tr = np.frombuffer(buf, np.float32, count=n * 31, offset=16).reshape(n, 31)
    out["tr_u"], out["tr_T"] = tr[:, 0].copy(), tr[:, 1:].copy()
    out["params"] = np.array([nused, maxdist, stepsize, wanted, shoot_step], np.float64)

    # BVH hit ids
    arays = aggregate_test_rays(3000, 11, hit_bound, hit_target)
    af = os.path.join(tmp, "arays.bin"); sceneio.write_rays(af, arays)
    run(scene_file, "--intersect", af, os.path.join(tmp, "hits.bin"))
    prim, t, occ = sceneio.read_hits(os.path.join(tmp, "hits.bin"))
    out["hit_rays"], out["hit_prim"], out["hit_t"], out["hit_occluded"] = arays, prim, t, occ
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def golden_synthetic_knn(tmp, scene_file):
    """Kernel-isolated k-NN / LPhoton golden on a uniform synthetic photon set."""
    n = 20000
    pos, wi, alpha = scenes.synthetic_photons(n)
    pho = os.path.join(tmp, "syn.pho"); sceneio.write_photons(pho, pos, wi, alpha)
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1.05, 1.05, size=(600, 3)).astype(np.float32)
    pts[:16] = pos[:16]                      # queries sitting exactly on photons (d2 == 0)
    w = np.tile(np.array([[0, 0, -1]], np.float32), (len(pts), 1))
    qf = os.path.join(tmp, "synq.bin"); sceneio.write_queries(qf, pts, w)
    out = dict(pos=pos, wi=wi, q_pts=pts)
    ops = ["--load-photons", pho]
    cases = [(50, 0.25 ** 2), (8, 0.05 ** 2), (300, 0.5 ** 2), (64, 0.08 ** 2), (1, 1.0)]
    for k, r2 in cases:
        ops += ["--knn", qf, k, repr(float(np.float32(r2))), os.path.join(tmp, "synknn_%d.bin" % k)]
    run(scene_file, *ops)
    for k, r2 in cases:
        nf, idx, d2 = sceneio.read_knn(os.path.join(tmp, "synknn_%d.bin" % k))
        out["knn%d_nfound" % k], out["knn%d_idx" % k], out["knn%d_d2" % k] = nf, idx, d2
        out["knn%d_r2" % k] = np.array([r2], np.float32)
    np.savez_compressed(os.path.join(HERE, "synthetic_knn.npz"), **out)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as tmp:
        homog = os.path.join(tmp, "cornell_homog.pbrt")
        open(homog, "w").write(scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 6000))
        golden_for_scene(tmp, "cornell_homog", homog, 50, 0.25, 0.05, 96, 64, [(50, 0.25 ** 2), (16, 0.1 ** 2)], wanted=6000)
        golden_synthetic_knn(tmp, homog)

        n = 32
        dens = scenes.blob_density(n)
        grid = os.path.join(tmp, "cornell_grid.pbrt")
        open(grid, "w").write(scenes.cornell_pbrt(scenes.grid_volume_text(n, dens), 2500, stepsize=0.0625, nused=50,
                                                  maxdist=0.25, shoot_step=0.05))
        golden_for_scene(tmp, "cornell_grid32", grid, 50, 0.25, 0.0625, 64, 64, [(50, 0.25 ** 2)], wanted=2500)
        g = sceneio.read_scene(os.path.join(HERE, "cornell_grid32.scn"))
        assert np.array_equal(g.density, dens), "density grid did not survive the text round trip"
        project_goldens(tmp)
        surface_goldens(tmp)
        sphere_goldens(tmp)
        exponential_goldens(tmp)
        volint_goldens(tmp)
        aggregate_goldens(tmp)
        area_light_goldens(tmp)
        underflow_goldens(tmp)


def project_goldens(tmp):
    """BASELINE configs 1 and 4 (the reference project's own scenes, parameters in scenes.py).
    Kernel-level goldens use the scenes with final gathering off and no caustic map, so that the reference's single MT19937
    stream is consumed by the volume path alone and the oracle can replay it; the end-to-end images use the shipped settings
    (config 1) / reduced counts (config 4) and are rendered by the unmodified reference binary."""
    # config 1: rainbow medium [-10,0,-5]-[5,5,5] under Translate(0,-.5,3.5), distant light, camera at the origin looking +z
    vol = os.path.join(tmp, "rainbow_vol.pbrt")
    open(vol, "w").write(scenes.volumescene_pbrt(nphotons=3000, caustic=0, finalgather=False, xres=64, yres=64))
    golden_for_scene(tmp, "rainbow_vol", vol, 50, 0.5, 0.15, 64, 48, [(50, 0.5 ** 2)], camera=dict(fov_deg=70.0, eye=(0, 0, 0), look=(0, 0, 1)),
                     box=((-10, -0.5, -1.5), (5, 4.5, 8.5)), wanted=3000, shoot_step=0.1, hit_bound=(-6.0, 9.0))
    # config 4 (reduced): glass wedge with dispersion (Vn 2.75), spot + point light, homogeneous medium
    prism = os.path.join(tmp, "prism_small.pbrt")
    open(prism, "w").write(scenes.prism_pbrt(nphotons=4000, caustic=0, nused=100, xres=64, yres=64, spp=1))
    golden_for_scene(tmp, "prism_small", prism, 100, 0.4, 0.05, 48, 48, [(100, 0.4 ** 2)],
                     camera=dict(fov_deg=70.0, eye=(0, 0, 0), look=(0, 0.0872, 0.9962)), box=((-10, -10.5, -6.5), (5, 4.5, 8.5)),
                     wanted=4000, shoot_step=0.1, hit_bound=(-6.0, 9.0), hit_target=((0.1, 0.85, 3.5), (0.9, 0.8, 0.1)), q_near_photons=True)
    # end-to-end reference images
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref")
    surf = scenes.cornell_surf_pbrt(nphotons=20000, caustic=5000, indirect=10000, finalgather=True, fgsamples=16, xres=72, yres=72,
                                    outfile="cornell_surf_e2e.pfm").replace('"integer pixelsamples" [1]', '"integer pixelsamples" [4]')
    for name, text in (("config1_volumescene", scenes.volumescene_pbrt(outfile="config1_volumescene.pfm", xres=150, yres=150)),
                       ("config4_prism", scenes.prism_pbrt(nphotons=20000, nused=200, xres=96, yres=96, spp=8, outfile="config4_prism.pfm")),
                       ("cornell_surf_e2e", surf),
                       ("sphere_e2e", scenes.sphere_pbrt(nphotons=100000, caustic=5000, finalgather=True, fgsamples=16, surf_nused=100, nused=100,
                                                         xres=96, yres=96, spp=4, outfile="sphere_e2e.pfm"))):
        f = os.path.join(tmp, name + ".pbrt"); open(f, "w").write(text)
        open(os.path.join(ROOT, "tests", "scenes", name + ".pbrt"), "w").write(text)
        subprocess.check_call([ref_bin, "--ncores", "1", "--quiet", f], cwd=tmp)
        np.save(os.path.join(HERE, name + "_ref.npy"), read_pfm(os.path.join(tmp, name + ".pfm")).astype(np.float16))
    exr_check(tmp, ref_bin)


def exr_check(tmp, ref_bin):
    """The reference's EXR writer (core/imageio.cpp:171-197, compiled in by oracle/Makefile from the vendored OpenEXR): config 1's
    scene with an .exr file name decodes to exactly the fp16 golden made from its .pfm -- so config1_volumescene_ref.npy IS the
    reference's EXR image and the drop-in's .exr is compared with it (tests/test_dropin_render.py)."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    f = os.path.join(tmp, "config1_volumescene_exr.pbrt")
    open(f, "w").write(scenes.volumescene_pbrt(outfile="config1_volumescene.exr", xres=150, yres=150))
    subprocess.check_call([ref_bin, "--ncores", "1", "--quiet", f], cwd=tmp)
    bgra = cv2.imread(os.path.join(tmp, "config1_volumescene.exr"), cv2.IMREAD_UNCHANGED)
    gold = np.load(os.path.join(HERE, "config1_volumescene_ref.npy")).astype(np.float32)
    assert bgra.shape == (150, 150, 4) and np.array_equal(bgra[..., [2, 1, 0]], gold) and np.all(bgra[..., 3] == 1.0)


def sphere_goldens(tmp):
    """Sphere primitives (shapes/sphere.cpp).  sphere_glass = the parameters of the reference project's
    projectScene/scene.pbrt (glass ball in a homogeneous medium, spot + point light) at reduced counts; sphere_disp = the same with a
    rotated, non-uniformly scaled, PARTIAL sphere (zmin/zmax/phimax) of dispersive, reflecting glass."""
    cam = dict(fov_deg=70.0, eye=(0, 0, 0), look=(0, 0.0872, 0.9962))
    box = ((-11, -1, -1.5), (4, 4, 8.5))
    target = ((-1.0, 1.0, 3.5), (0.7, 0.7, 0.7))
    f = os.path.join(tmp, "sphere_glass.pbrt")
    open(f, "w").write(scenes.sphere_pbrt(nphotons=4000, caustic=0, finalgather=False, nused=100, xres=64, yres=64, spp=1))
    golden_for_scene(tmp, "sphere_glass", f, 100, 0.5, 0.05, 48, 48, [(100, 0.5 ** 2)], camera=cam, box=box, wanted=4000, shoot_step=0.1,
                     hit_bound=(-6.0, 9.0), hit_target=target, q_near_photons=True)
    f = os.path.join(tmp, "sphere_disp.pbrt")
    open(f, "w").write(scenes.sphere_pbrt(nphotons=4000, caustic=0, finalgather=False, nused=100, xres=64, yres=64, spp=1, vn=4.0, kr=0.5,
                                          sphere_xform="Rotate 30 1 0 0\nScale 1 0.8 1.2\n",
                                          sphere_params=' "float zmin" [-0.4] "float zmax" [0.5] "float phimax" [300]'))
    golden_for_scene(tmp, "sphere_disp", f, 100, 0.5, 0.05, 48, 48, [(100, 0.5 ** 2)], camera=cam, box=box, wanted=4000, shoot_step=0.1,
                     hit_bound=(-6.0, 9.0), hit_target=target, q_near_photons=True)


def exponential_goldens(tmp):
    """ExponentialDensity medium (volumes/exponential.h) in the Cornell box."""
    f = os.path.join(tmp, "cornell_exp.pbrt")
    open(f, "w").write(scenes.cornell_pbrt(scenes.EXP_VOLUME, 3000, stepsize=0.05, nused=50, maxdist=0.25, shoot_step=0.05))
    golden_for_scene(tmp, "cornell_exp", f, 50, 0.25, 0.05, 64, 64, [(50, 0.25 ** 2)], wanted=3000)


def volint_goldens(tmp):
    """SingleScatteringIntegrator::Li and EmissionIntegrator::Li of the unmodified reference (ref_harness --vli)
    on three media; one RNG(seed + i) per ray so the oracle's MT mode can replay every draw."""
    out = {}
    cases = dict(scenes.VOLINT_MEDIA)
    cases["volint_grid"] = (scenes.volint_grid_volume(32), False, 0.0625)
    rays = scenes.camera_rays(64, 64)
    rng = np.random.default_rng(19)
    rays = rays[np.sort(rng.choice(len(rays), size=160, replace=False))]
    rays["u_scatter"] = rng.random(len(rays)).astype(np.float32)
    rays["u_scatter"][:8] = np.float32(0.5)
    rays["maxt"][8:24] = rng.uniform(2.6, 4.2, size=16).astype(np.float32)     # rays that end inside the medium
    out["rays"] = rays
    rf = os.path.join(tmp, "volint_rays.bin"); sceneio.write_rays(rf, rays)
    for name, (vol, second_light, stepsize) in cases.items():
        for kind in ("single", "emission"):
            f = os.path.join(tmp, "%s_%s.pbrt" % (name, kind))
            open(f, "w").write(scenes.volint_pbrt(kind, vol, stepsize=stepsize, second_light=second_light))
            ops = ["--vli", rf, 4000, os.path.join(tmp, "vli.bin")]
            if kind == "single":
                ops = ["--export-scene", os.path.join(HERE, name + ".scn")] + ops
            run(f, *ops)
            li = sceneio.read_spectra(os.path.join(tmp, "vli.bin"), b"PVLI0001", per=2)
            out["%s_%s_L" % (name, kind)], out["%s_%s_T" % (name, kind)] = li[:, 0], li[:, 1]
            print("  %s/%s: mean L %.4g, min T.y %.3g" % (name, kind, li[:, 0].mean(), li[:, 1].mean(axis=1).min()))
        out[name + "_stepsize"] = np.array([stepsize], np.float32)
    # a scene off the device path except for its medium (disk-shaped area light): the full export must refuse it, the
    # medium-only export (what the drop-in falls back to for "emission") carries all that EmissionIntegrator::Li reads
    f = os.path.join(tmp, "volint_offpath.pbrt"); open(f, "w").write(scenes.volint_offpath_pbrt())
    assert subprocess.run([HARNESS, f, "--export-scene", os.path.join(tmp, "refused.scn")], capture_output=True).returncode == 4
    run(f, "--export-medium", os.path.join(HERE, "volint_offpath_medium.scn"), "--vli", rf, 4000, os.path.join(tmp, "vli.bin"))
    li = sceneio.read_spectra(os.path.join(tmp, "vli.bin"), b"PVLI0001", per=2)
    out["volint_offpath_emission_L"], out["volint_offpath_emission_T"] = li[:, 0], li[:, 1]
    # edge-case rays (CPU oracle pin only): origins inside and outside the medium, non-unit and axis-aligned directions, finite
    # extents, mint > 0, rays that only graze or miss the box, zero-length medium intervals
    erays = aggregate_test_rays(96, 29, (-1.6, 1.6))
    erays["u_scatter"] = np.random.default_rng(31).random(len(erays)).astype(np.float32)
    erays["o"][:4] = np.float32([[-1, 0.2, -3], [1, -0.3, -3], [0.5, 1, -3], [0, 0, -1]]); erays["d"][:4] = np.float32([0, 0, 1])   # along faces / from a face
    erays["mint"][:4] = 0; erays["maxt"][:4] = np.float32([np.inf, np.inf, np.inf, 0.0])
    out["edge_rays"] = erays
    ef = os.path.join(tmp, "volint_edge_rays.bin"); sceneio.write_rays(ef, erays)
    for name in ("volint_homog", "volint_grid"):
        for kind in ("single", "emission"):
            run(os.path.join(tmp, "%s_%s.pbrt" % (name, kind)), "--vli", ef, 5000, os.path.join(tmp, "vli.bin"))
            li = sceneio.read_spectra(os.path.join(tmp, "vli.bin"), b"PVLI0001", per=2)
            out["edge_%s_%s_L" % (name, kind)], out["edge_%s_%s_T" % (name, kind)] = li[:, 0], li[:, 1]
    # AggregateVolume: two overlapping Volume statements, point + spot light; the scene is exported once per region
    agg = scenes.aggregate_volumes(32)
    for kind in ("single", "emission"):
        f = os.path.join(tmp, "volint_agg_%s.pbrt" % kind)
        open(f, "w").write(scenes.volint_pbrt(kind, agg, stepsize=0.0625, second_light=True))
        ops = ["--vli", rf, 4000, os.path.join(tmp, "vli.bin"), "--vli", ef, 5000, os.path.join(tmp, "vli_e.bin")]
        if kind == "single":
            ops = ["--export-regions", os.path.join(HERE, "volint_agg"), "--transmittance", rf, 77, os.path.join(tmp, "tr.bin")] + ops
        run(f, *ops)
        li = sceneio.read_spectra(os.path.join(tmp, "vli.bin"), b"PVLI0001", per=2)
        out["volint_agg_%s_L" % kind], out["volint_agg_%s_T" % kind] = li[:, 0], li[:, 1]
        li = sceneio.read_spectra(os.path.join(tmp, "vli_e.bin"), b"PVLI0001", per=2)
        out["edge_volint_agg_%s_L" % kind], out["edge_volint_agg_%s_T" % kind] = li[:, 0], li[:, 1]
    tr = np.frombuffer(open(os.path.join(tmp, "tr.bin"), "rb").read(), np.float32, count=len(rays) * 31, offset=16).reshape(len(rays), 31)
    out["volint_agg_tr_u"], out["volint_agg_tr_T"] = tr[:, 0].copy(), tr[:, 1:].copy()
    out["volint_agg_stepsize"] = np.array([0.0625], np.float32)
    # a DiffuseAreaLight next to the point light: scene with a placeholder in the area light's slot + the light's data on the side
    f = os.path.join(tmp, "volint_area.pbrt"); open(f, "w").write(scenes.volint_area_pbrt("single"))
    run(f, "--export-area-lights", os.path.join(HERE, "volint_area.scn"), os.path.join(HERE, "volint_area.lights"),
        "--vli", rf, 4000, os.path.join(tmp, "vli.bin"), "--vli", ef, 5000, os.path.join(tmp, "vli_e.bin"))
    li = sceneio.read_spectra(os.path.join(tmp, "vli.bin"), b"PVLI0001", per=2)
    out["volint_area_single_L"], out["volint_area_single_T"] = li[:, 0], li[:, 1]
    li = sceneio.read_spectra(os.path.join(tmp, "vli_e.bin"), b"PVLI0001", per=2)
    out["edge_volint_area_single_L"], out["edge_volint_area_single_T"] = li[:, 0], li[:, 1]
    out["mt_seed"] = np.array([4000], np.uint32)
    np.savez_compressed(os.path.join(HERE, "volint.npz"), **out)
    # end-to-end images of the unmodified reference binary (glass wedge: specular bounces call the volume integrator per ray)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref")
    for name, kind, vol, stepsize in (("volint_single_e2e", "single", scenes.VOLINT_MEDIA["volint_homog"][0], 0.05),
                                      ("volint_emission_e2e", "emission", scenes.volint_grid_volume(32), 0.0625)):
        text = scenes.volint_e2e_pbrt(kind, vol, stepsize=stepsize, outfile=name + ".pfm")
        f = os.path.join(tmp, name + ".pbrt"); open(f, "w").write(text)
        open(os.path.join(ROOT, "tests", "scenes", name + ".pbrt"), "w").write(text)
        subprocess.check_call([ref_bin, "--ncores", "1", "--quiet", f], cwd=tmp)
        img = read_pfm(os.path.join(tmp, name + ".pfm"))
        np.save(os.path.join(HERE, name + "_ref.npy"), img.astype(np.float16))
        subprocess.check_call([ref_bin, "--ncores", "3", "--quiet", f], cwd=tmp)          # another task count = other RNG seeds: the noise floor
        img2 = read_pfm(os.path.join(tmp, name + ".pfm"))
        lum = lambda a: 0.2126 * a[..., 0] + 0.7152 * a[..., 1] + 0.0722 * a[..., 2]
        h = img.shape[0] // 6 * 6
        bm = lambda a: lum(a)[:h, :h].reshape(h // 6, 6, h // 6, 6).mean(axis=(1, 3))
        b1, b2 = bm(img), bm(img2); lit = b1 > 0.05 * b1.mean()
        print("  %s: mean lum %.4g; two reference runs differ by %.3f%% (mean), %.3f%% (block MRE)" % (
            name, lum(img).mean(), 100 * abs(lum(img).mean() - lum(img2).mean()) / lum(img).mean(), 100 * (np.abs(b1 - b2)[lit] / b1[lit]).mean()))


def aggregate_goldens(tmp):
    """The photon-volume path under an AggregateVolume (two overlapping Volume statements, core/volume.cpp:178-261): the reference's
    photon list, k-NN sets, LPhoton, Li, transmittance, hits -- for the oracle now, for the device path's aggregate next."""
    f = os.path.join(tmp, "cornell_agg.pbrt")
    open(f, "w").write(scenes.cornell_pbrt(scenes.aggregate_volumes(32), 2500, stepsize=0.0625, nused=50, maxdist=0.25, shoot_step=0.05))
    golden_for_scene(tmp, "cornell_agg", f, 50, 0.25, 0.0625, 64, 64, [(50, 0.25 ** 2)], wanted=2500, regions=True)


def area_light_goldens(tmp):
    """A DiffuseAreaLight next to the point light in the photon-volume scene: the reference's photon list of one task.  Photons now
    start on the light's quad (area-CDF shape choice, Triangle::Sample, hemisphere direction, pdf = nShapes / sumArea / 2 pi)."""
    f = os.path.join(tmp, "cornell_area.pbrt")
    open(f, "w").write(scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 3000).replace("WorldEnd", scenes.AREA_QUAD + "\nWorldEnd"))
    pho = os.path.join(tmp, "cornell_area.pho"); stats = os.path.join(tmp, "cornell_area.json")
    run(f, "--export-area-lights", os.path.join(HERE, "cornell_area.scn"), os.path.join(HERE, "cornell_area.lights"),
        "--shoot", "--dump-photons", pho, "--stats", stats)
    pos, wi, alpha = sceneio.read_photons(pho)
    st = json.load(open(stats))
    print("  cornell_area: %d photons from %d paths" % (len(pos), st["nshot"]))
    # PhotonVolumeIntegrator::Li on that photon list: the direct term samples the area light too
    rays = scenes.camera_rays(64, 64)
    rng = np.random.default_rng(41)
    rays = rays[np.sort(rng.choice(len(rays), size=64, replace=False))]
    rays["u_scatter"] = rng.random(len(rays)).astype(np.float32)
    rf = os.path.join(tmp, "area_rays.bin"); sceneio.write_rays(rf, rays)
    run(f, "--load-photons", pho, "--li", rf, 1000, os.path.join(tmp, "area_li.bin"))
    li = sceneio.read_spectra(os.path.join(tmp, "area_li.bin"), b"PVLI0001", per=2)
    np.savez_compressed(os.path.join(HERE, "cornell_area.npz"), shot_pos=pos, shot_wi=wi, shot_alpha=alpha, nshot=np.array([st["nshot"]], np.uint64),
                        params=np.array([50, 0.25, 0.05, 3000, 0.05], np.float64), li_rays=rays, li_L=li[:, 0], li_T=li[:, 1])


def underflow_goldens(tmp, seed=113):
    """The scene on which the randomised cross-check (fuzz_oracle.py, seed 113) separated `Spectrum::lambda as path state` from
    `lambda inferred from the bins`: a dispersive glass wedge in an exponential medium dense enough for photon weights of 1e-36 to
    underflow to zero between two glass faces.  Stored: the flattened scene and the photon list of one reference task."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_oracle", os.path.join(HERE, "fuzz_oracle.py"))
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    d = fz.scene_for_seed(seed)
    assert d["integ"] == "photonvolume" and d["variant"] == "glass"
    f = os.path.join(tmp, "underflow_glass.pbrt"); open(f, "w").write(d["text"])
    pho = os.path.join(tmp, "u.pho"); stats = os.path.join(tmp, "u.json")
    run(f, "--ncores", 1, "--export-scene", os.path.join(HERE, "underflow_glass.scn"), "--shoot", "--dump-photons", pho, "--stats", stats)
    pos, wi, alpha = sceneio.read_photons(pho); st = json.load(open(stats))
    np.savez_compressed(os.path.join(HERE, "underflow_glass.npz"), shot_pos=pos, shot_wi=wi, shot_alpha=alpha, nshot=np.array([st["nshot"]], np.uint64),
                        params=np.array([d["wanted"], d["shoot_step"], d["stepsize"]], np.float64))
    print("underflow_glass: %d photons, %d black, nshot %d" % (len(pos), int((alpha.max(axis=1) == 0).sum()), st["nshot"]))


def read_radiance(fn):
    """PVRADP01 (oracle/ref_harness.cpp --dump-maps): p[3] n[3] Lo[30] rho_r[30] rho_t[30] per radiance photon."""
    buf = open(fn, "rb").read()
    assert buf[:8] == b"PVRADP01"
    n = int(np.frombuffer(buf, np.uint64, 1, 8)[0])
    r = np.frombuffer(buf, np.float32, 96 * n, 16).reshape(n, 96)
    return dict(pos=r[:, 0:3].copy(), n=r[:, 3:6].copy(), Lo=r[:, 6:36].copy(), rho_r=r[:, 36:66].copy(), rho_t=r[:, 66:96].copy())


# name -> (scene text, (volume, caustic, indirect) wanted, final gather, shooter step, integrator step)
SURFACE_CASES = {
    # every map on: glass wedge (no dispersion) under the light, matte walls, homogeneous medium
    "cornell_surf": (lambda: scenes.cornell_surf_pbrt(nphotons=1500, caustic=800, indirect=2000), (1500, 800, 2000), True, 0.05, 0.05),
    # dispersive wedge (Vn 2), caustic map finishes last, indirect first
    "cornell_surf_disp": (lambda: scenes.cornell_surf_pbrt(vn=2.0, nphotons=2500, caustic=1500, indirect=600), (2500, 1500, 600), True, 0.05, 0.05),
    # BASELINE config 1 with its shipped surface-integrator settings (caustic map + final gathering), reduced counts
    "rainbow_surf": (lambda: scenes.volumescene_pbrt(nphotons=1500, caustic=1000, finalgather=True, xres=64, yres=64), (1500, 1000, 0), True, 0.1, 0.15),
}


def surface_goldens(tmp):
    """the reference's own photon lists of ALL maps (one task => deterministic) and its radiance photons."""
    for name, (text, wanted, fg, shoot_step, istep) in SURFACE_CASES.items():
        f = os.path.join(tmp, name + ".pbrt"); open(f, "w").write(text())
        prefix = os.path.join(tmp, name)
        run(f, "--export-scene", os.path.join(HERE, name + ".scn"), "--shoot", "--dump-photons", prefix + ".volume", "--dump-maps", prefix,
            "--stats", prefix + ".json")
        st = json.load(open(prefix + ".json"))
        out = {}
        for k in ("volume", "caustic", "indirect", "direct"):
            pos, wi, alpha = sceneio.read_photons(prefix + "." + k)
            out[k + "_pos"], out[k + "_wi"], out[k + "_alpha"] = pos, wi, alpha
        rad = read_radiance(prefix + ".radiance")
        assert not rad["rho_t"].any()
        out["rad_pos"], out["rad_n"], out["rad_Lo"], out["rad_rho_r"] = rad["pos"], rad["n"], rad["Lo"], rad["rho_r"]
        out["counts"] = np.array([st["nshot"], st["caustic_paths"], st["indirect_paths"], st["direct_paths"], st["volume_paths"]], np.uint64)
        # the surface integrator's photon lookups through the reference's own code: LPhoton (caustic / indirect map) and the
        # radiance-photon lookup of final gathering, at radiance-photon sites and at points jittered off them
        rng = np.random.default_rng(23)
        pick = rng.choice(len(rad["pos"]), size=min(200, len(rad["pos"])), replace=False)
        qp = np.concatenate([rad["pos"][pick], rad["pos"][pick] + rng.uniform(-0.06, 0.06, size=(len(pick), 3))]).astype(np.float32)
        qn = np.concatenate([rad["n"][pick], rad["n"][pick] * np.where(rng.random(len(pick)) < 0.3, -1.0, 1.0)[:, None]]).astype(np.float32)
        qf = os.path.join(tmp, name + "_sq.bin"); sceneio.write_queries(qf, qp, qn)
        run(f, "--shoot", "--surface-lphoton", "caustic", qf, prefix + ".slc", "--surface-lphoton", "indirect", qf, prefix + ".sli",
            "--radiance-nearest", qf, prefix + ".rn")
        out["sq_pts"], out["sq_n"] = qp, qn
        for key, ext in (("caustic", ".slc"), ("indirect", ".sli")):
            v = sceneio.read_spectra(prefix + ext, b"PVSLPH01", per=2)
            out["slp_%s_Lr_pi" % key], out["slp_%s_Lt_pi" % key] = v[:, 0], v[:, 1]
        buf = open(prefix + ".rn", "rb").read(); assert buf[:8] == b"PVRADN01"
        rec = np.frombuffer(buf, np.uint32, count=len(qp) * 32, offset=16).reshape(len(qp), 32)
        out["radn_idx"] = rec[:, 0].copy(); out["radn_d2"] = rec[:, 1].copy().view(np.float32); out["radn_Lo"] = rec[:, 2:].copy().view(np.float32)
        out["params"] = np.array([wanted[0], wanted[1], wanted[2], int(fg), shoot_step, istep, st["nlookup"], st["maxdist2"]], np.float64)
        print("  %s: nshot %d, volume %d caustic %d indirect %d direct %d radiance %d" % (
            name, st["nshot"], len(out["volume_pos"]), len(out["caustic_pos"]), len(out["indirect_pos"]), len(out["direct_pos"]), len(rad["pos"])))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def read_pfm(path):
    with open(path, "rb") as f:
        kind = f.readline().strip()
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, 3 if kind == b"PF" else 1)
    return data[::-1].copy()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "exponential":
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp_:
            exponential_goldens(tmp_)
    elif len(sys.argv) > 1 and sys.argv[1] == "area":             # only the area-light photon list
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp_:
            area_light_goldens(tmp_)
    elif len(sys.argv) > 1 and sys.argv[1] == "aggregate":        # only the AggregateVolume goldens of the photon-volume path
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp_:
            aggregate_goldens(tmp_)
    elif len(sys.argv) > 1 and sys.argv[1] == "volint":           # only the single / emission integrator goldens
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp_:
            volint_goldens(tmp_)
    elif len(sys.argv) > 1 and sys.argv[1] == "sphere":           # only the sphere-scene goldens
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp_:
            sphere_goldens(tmp_)
    elif len(sys.argv) > 1 and sys.argv[1] == "surface":          # only the surface-map goldens
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp_:
            surface_goldens(tmp_)
    elif len(sys.argv) > 1 and sys.argv[1] == "project":          # only the config-1 / config-4 goldens
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
        with tempfile.TemporaryDirectory() as tmp_:
            project_goldens(tmp_)
    else:
        main()
