"""GPU (-m gpu): VolumeIntegrator "single" / "emission" on the device (pv_volume_li, csrc/pv_volint.cu) through
the C ABI, against (a) what the real reference returned (tests/golden/volint.npz, ref_harness --vli) wherever the value does not
depend on a random draw and (b) the pinned CPU oracle on the same keyed Philox stream everywhere else.
Bar: radiance AND transmittance within 1e-4 relative (north_star's per-ray tolerance; the transmittance of these two integrators
is a product over all march steps, so it carries the libm-vs-CUDA expf difference of every step), identical zero patterns
(the Russian-roulette decisions must agree)."""
import os
import numpy as np
import pytest
import oracle_lib as O
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
RTOL = 1e-4
KINDS = {"single": O.SINGLE, "emission": O.EMISSION}


def relerr(a, b, floor=1e-30):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)




def volint_scene(pkg, golden, name):
    g, _ = golden("volint")
    return g, pkg.sceneio.read_scene(os.path.join(GOLDEN, name + ".scn")), float(g[name + "_stepsize"][0])


def check(L, T, oL, oT, need_light=True):
    assert np.array_equal(T == 0, oT == 0)                      # same roulette outcomes
    assert relerr(T, oT)[oT > 0].max() < RTOL
    assert np.array_equal(L == 0, oL == 0)
    m = oL > 0
    if need_light:
        assert m.any()
    if m.any():
        assert relerr(L, oL)[m].max() < RTOL


@pytest.mark.parametrize("name,kind", [("volint_homog", "emission"), ("volint_dense", "single"), ("volint_dense", "emission")])
def test_volume_li_vs_reference_where_no_draw_matters(golden, pkg, pv_factory, name, kind):
    """Homogeneous media ignore the tau offsets; with one light (or none used) the only draw left is the Russian roulette
    (volint_homog under "single" has two lights, so its light choice is a draw: that case is checked against the oracle below).
    On the rays where the oracle under the CUDA path's own Philox stream reproduces the reference's MT19937-driven value bit
    for bit (no roulette, or the same roulette outcomes) the CUDA result is compared directly with what the reference binary
    returned."""
    g, scene, stepsize = volint_scene(pkg, golden, name)
    rays = g["rays"]
    refL, refT = g["%s_%s_L" % (name, kind)], g["%s_%s_T" % (name, kind)]
    aL, aT, _ = O.volume_li(scene, rays, stepsize, KINDS[kind], seed=99, rng_mode=O.PHILOX)
    fixed = (aL.view(np.uint32) == refL.view(np.uint32)).all(axis=1) & (aT.view(np.uint32) == refT.view(np.uint32)).all(axis=1)
    if name == "volint_homog" and kind == "emission":
        assert fixed.all()
    assert fixed.sum() >= 40
    pv = pv_factory(stepsize=stepsize, seed=99)
    pv.set_scene(scene)
    L, T = pv.VolumeLi(kind, rays)
    check(L[fixed], T[fixed], refL[fixed], refT[fixed])


@pytest.mark.parametrize("kind", ["single", "emission"])
@pytest.mark.parametrize("name", ["volint_homog", "volint_dense", "volint_grid"])
def test_volume_li_vs_oracle_same_philox_stream(golden, pkg, pv_factory, name, kind):
    g, scene, stepsize = volint_scene(pkg, golden, name)
    rays = g["rays"]
    pv = pv_factory(stepsize=stepsize, seed=0xC0FFEE)
    pv.set_scene(scene)
    L, T = pv.VolumeLi(kind, rays, ray_index_base=700)
    oL, oT, _ = O.volume_li(scene, rays, stepsize, KINDS[kind], seed=0xC0FFEE, ray_index_base=700)
    check(L, T, oL, oT)
    if name == "volint_dense":
        assert (oT == 0).all(axis=1).any() and ((oT > 0) & (oT < 1e-2)).any()     # roulette both ended and spared marches
    # the two schedules (one warp per ray / one thread per ray) run the same arithmetic: bit-identical
    A = pkg._abi
    Lw, Tw = pv.VolumeLi(kind, rays, ray_index_base=700, flags=A.VOLINT_WARP_PER_RAY)
    Lt, Tt = pv.VolumeLi(kind, rays, ray_index_base=700, flags=A.VOLINT_THREAD_PER_RAY)
    assert np.array_equal(L, Lw) and np.array_equal(T, Tw)            # the default for 160 rays
    assert np.array_equal(Lt.view(np.uint32), Lw.view(np.uint32)) and np.array_equal(Tt.view(np.uint32), Tw.view(np.uint32))
    # sharding by ray_index_base reproduces the one-call result bit for bit
    L2, T2 = pv.VolumeLi(kind, rays[50:120], ray_index_base=750)
    assert np.array_equal(L2, L[50:120]) and np.array_equal(T2, T[50:120])


@pytest.mark.parametrize("name", ["rainbow_vol", "cornell_exp", "sphere_glass", "prism_small"])
def test_single_li_on_the_path_scenes_vs_oracle(golden, pv_factory, name):
    """The scenes of the photon-volume goldens under the single-scattering integrator: rainbow medium (RainbowVolume::p is the
    homogeneous medium's PhaseHG), exponential medium, sphere primitives in the shadow rays, spot + point light."""
    g, scene = golden(name)
    stepsize = float(g["params"][2])
    rays = g["li_rays"]
    pv = pv_factory(stepsize=stepsize, seed=31337)
    pv.set_scene(scene)
    L, T = pv.VolumeLi("single", rays, ray_index_base=5)
    oL, oT, ost = O.volume_li(scene, rays, stepsize, O.SINGLE, seed=31337, ray_index_base=5)
    assert ost.shadow_rays > 0
    check(L, T, oL, oT)
    Lt, Tt = pv.VolumeLi("single", rays, ray_index_base=5, flags=8)          # PV_VOLINT_THREAD_PER_RAY
    assert np.array_equal(Lt.view(np.uint32), L.view(np.uint32)) and np.array_equal(Tt.view(np.uint32), T.view(np.uint32))
    Le, Te = pv.VolumeLi("emission", rays, ray_index_base=5)
    oLe, oTe, _ = O.volume_li(scene, rays, stepsize, O.EMISSION, seed=31337, ray_index_base=5)
    check(Le, Te, oLe, oTe, need_light=False)


def test_volume_li_edge_cases(golden, pkg, pv_factory):
    g, scene, stepsize = volint_scene(pkg, golden, "volint_homog")
    pv = pv_factory(stepsize=stepsize)
    with pytest.raises(pkg.PVError):
        pv.VolumeLi("single", g["rays"][:4])                    # no scene yet
    pv.set_scene(scene)
    rays = pkg.sceneio.make_rays(np.array([[0, 5, -5], [0, 0, -3]], np.float32), np.array([[0, 0, 1], [0, 0, 1]], np.float32))
    rays["maxt"][1] = 1.5                                        # stops before the box
    for kind in KINDS:
        L, T = pv.VolumeLi(kind, rays)
        assert (L == 0).all() and (T == 1).all()
        L, T = pv.VolumeLi(kind, rays[:0])
        assert L.shape == (0, 30)
    import ctypes as C
    prm = pv.gather_params(0, 0)
    out = np.zeros((2, 30), np.float32)
    rc = pv.lib.pv_volume_li(pv.ctx, C.c_int(7), rays.ctypes.data_as(C.c_void_p), C.c_uint64(2), C.byref(prm),
                             out.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == -1 and b"unknown integrator" in pv.lib.pv_last_error(pv.ctx)          # PV_EINVAL
    # the photon-volume path of the same context is untouched by the call
    assert pv.photon_count() == 0


@pytest.fixture(scope="module")
def full_frame(pkg):
    """BASELINE config 3's scene shape at full size: 256^3 density grid (made emitting), 1920x1080 camera rays, stepsize 2/64."""
    import importlib
    W = importlib.import_module("cs348b_pbrt_b200.workloads")
    cfg = W.CONFIGS["config3"]
    scene = W.load_scene(cfg)
    for b in range(pkg._abi.NSPEC):
        scene.medium.le[b] = 0.05 + 0.01 * b
    rays, _ = W.frame_rays(cfg)
    pv = pkg.PhotonVolume(device=0, stepsize=cfg["stepsize"], seed=348)
    pv.set_scene(scene)
    yield dict(cfg=cfg, scene=scene, rays=rays, pv=pv)
    pv.close()


@pytest.mark.parametrize("kind", ["single", "emission"])
def test_full_frame_volume_li_properties(full_frame, pkg, kind):
    """The oracle cannot march 2 M rays in seconds, so at full size parity rests on properties that do not depend on size:
    determinism; the two kernel schedules agree bit for bit (frames take the thread-per-ray one by default); sharding the frame
    by ray_index_base reproduces it bit for bit; the radiance is exactly linear in the emission and the light intensity (a
    power-of-two scale is exact in fp32) while the transmittance does not move; and a sample of the frame against the oracle."""
    A = pkg._abi
    pv, rays, scene, cfg = full_frame["pv"], full_frame["rays"], full_frame["scene"], full_frame["cfg"]
    n = len(rays)
    L1, T1 = pv.VolumeLi(kind, rays)
    L2, T2 = pv.VolumeLi(kind, rays)
    assert np.array_equal(L1, L2) and np.array_equal(T1, T2)
    assert np.isfinite(L1).all() and L1.max() > 0 and (T1 >= 0).all() and (T1 <= 1).all()
    Lw, Tw = pv.VolumeLi(kind, rays, flags=A.VOLINT_WARP_PER_RAY)
    assert np.array_equal(L1.view(np.uint32), Lw.view(np.uint32)) and np.array_equal(T1.view(np.uint32), Tw.view(np.uint32))
    h = n // 2 + 7                                          # not tile aligned
    La, Ta = pv.VolumeLi(kind, rays[:h], ray_index_base=0)
    Lb, Tb = pv.VolumeLi(kind, rays[h:], ray_index_base=h)
    assert np.array_equal(np.concatenate([La, Lb]), L1) and np.array_equal(np.concatenate([Ta, Tb]), T1)
    # linearity: emission and every light twice as bright
    try:
        for b in range(A.NSPEC):
            scene.medium.le[b] *= 2.0
            for l in scene.lights:
                l.intensity[b] *= 2.0
        pv.set_scene(scene)
        L3, T3 = pv.VolumeLi(kind, rays)
    finally:
        for b in range(A.NSPEC):
            scene.medium.le[b] *= 0.5
            for l in scene.lights:
                l.intensity[b] *= 0.5
        pv.set_scene(scene)
    assert np.array_equal(L3, 2.0 * L1) and np.array_equal(T3, T1)
    # a sample of the frame against the pinned oracle (same Philox stream: the sample is its own call on both sides)
    sub = np.ascontiguousarray(rays[::400])
    gL, gT = pv.VolumeLi(kind, sub, flags=A.VOLINT_THREAD_PER_RAY)
    oL, oT, _ = O.volume_li(scene, sub, cfg["stepsize"], KINDS[kind], seed=348)
    check(gL, gT, oL, oT)


def test_indexed_calls_give_every_ray_its_own_stream(golden, pkg, pv_factory):
    """pv_gather_indexed / pv_volume_li_indexed: ray i draws from the stream ray_index[i], so a ray's result is the one it gets
    alone under that index -- whatever rays share the call, in whatever order (what the drop-in's batching of secondary rays across
    render threads needs for reproducible renders).  Two lights + a grid medium: light choice, tau offsets and roulette all draw."""
    g, scene, stepsize = volint_scene(pkg, golden, "volint_homog")
    rays = g["rays"][:64]
    rng = np.random.default_rng(5)
    idx = rng.integers(1 << 40, 1 << 62, size=len(rays), dtype=np.uint64) | np.uint64(1 << 63)
    pv = pv_factory(stepsize=stepsize, seed=0xBEEF)
    pv.set_scene(scene)
    L, T = pv.LiIndexed(rays, idx, integrator="single")
    for i in (0, 7, 63):
        Li, Ti = pv.VolumeLi("single", rays[i:i + 1], ray_index_base=int(idx[i]))
        assert np.array_equal(Li[0], L[i]) and np.array_equal(Ti[0], T[i])
    perm = rng.permutation(len(rays))
    Lp, Tp = pv.LiIndexed(np.ascontiguousarray(rays[perm]), idx[perm], integrator="single")
    assert np.array_equal(Lp, L[perm]) and np.array_equal(Tp, T[perm])
    # the photon-volume integrator's indexed form, on a shot map
    g2, scene2 = golden("cornell_grid32")
    pv2 = pv_factory(stepsize=float(g2["params"][2]), nused=int(g2["params"][0]), maxdist=float(g2["params"][1]), seed=3)
    pv2.set_scene(scene2)
    pv2.set_photons(g2["shot_pos"], g2["shot_wi"], g2["shot_alpha"]); pv2.build()
    r2 = g2["li_rays"][:48]; i2 = idx[:48]
    L2, T2 = pv2.LiIndexed(r2, i2)
    for i in (0, 20, 47):
        Li, Ti = pv2.Li(r2[i:i + 1], ray_index_base=int(i2[i]))
        assert np.array_equal(Li[0], L2[i]) and np.array_equal(Ti[0], T2[i])
