"""GPU (-m gpu): DiffuseAreaLight on the device (PV_LIGHT_AREA, lights/diffuse.cpp:69-106 + ShapeSet core/light.cpp:114-172).
The oracle's area-light code is pinned on the reference bit for bit (tests/test_oracle_golden.py: single-scattering Li and the photon
list of a scene with an area light, MT stream); here the device is compared with that oracle on the same Philox streams:
  * the direct term of the single-scattering integrator and of PhotonVolumeIntegrator::Li (area-CDF triangle choice,
    Triangle::Sample, ShapeSet::Sample's re-intersection of every shape, ShapeSet::Pdf, the shortened visibility segment);
  * photon emission from the area light (point by area, direction uniform over the sphere flipped into the normal's hemisphere,
    pdf = ShapeSet::Pdf(point) / 2 pi, ray epsilon 1e-3), photons matched one to one by (path, deposit ordinal)."""
import os
import numpy as np
import pytest
import oracle_lib as O
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def relerr(a, b, floor=1e-30):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def area_scene(pkg, name):
    return pkg.sceneio.attach_area_lights(pkg.sceneio.read_scene(os.path.join(GOLDEN, name + ".scn")), os.path.join(GOLDEN, name + ".lights"))


def test_single_scattering_li_with_an_area_light_vs_oracle(golden, pkg, pv_factory):
    g, _ = golden("volint")
    scene = area_scene(pkg, "volint_area")
    assert [l.type for l in scene.lights] == [pkg._abi.LIGHT_POINT, pkg._abi.LIGHT_AREA]
    pv = pv_factory(stepsize=0.05, seed=0xA2EA)
    pv.set_scene(scene)
    for rays, base in ((g["rays"], 11), (g["edge_rays"], 900)):
        L, T = pv.VolumeLi("single", rays, ray_index_base=base)
        oL, oT, st = O.volume_li(scene, rays, 0.05, O.SINGLE, seed=0xA2EA, ray_index_base=base)
        assert st.shadow_rays > 0
        # (edge rays: a sample point in the plane of the light, or the sampled point itself, makes ShapeSet::Pdf 0/0 -- the
        # reference's own arithmetic, reproduced by the oracle under either stream; such rays are compared for being the same rays)
        fin = np.isfinite(oL).all(axis=1)
        assert fin.sum() >= 0.9 * len(rays) and np.array_equal(np.isfinite(L).all(axis=1), fin)
        lit = (oL > 1e-6 * oL[fin].max()) & fin[:, None]
        assert lit.any() and relerr(L, oL)[lit].max() < 1e-4 and np.abs(T - oT)[fin].max() < 1e-5
    # the area light contributes: without it (its slot dark) the radiance is visibly lower
    dark = pkg.sceneio.read_scene(os.path.join(GOLDEN, "volint_area.scn"))
    dL, _, _ = O.volume_li(dark, g["rays"], 0.05, O.SINGLE, seed=0xA2EA, ray_index_base=11)
    L, _ = pv.VolumeLi("single", g["rays"], ray_index_base=11)
    assert L.sum() > 1.02 * dL.sum() and (np.abs(L - dL) > 1e-3 * dL.max()).any(axis=1).mean() > 0.3


def test_shooter_and_gather_with_an_area_light_vs_oracle(golden, pkg, pv_factory):
    g, _ = golden("cornell_area")
    scene = area_scene(pkg, "cornell_area")
    nused, maxdist, istep, wanted, sstep = int(g["params"][0]), float(g["params"][1]), float(g["params"][2]), int(g["params"][3]), float(g["params"][4])
    pv = pv_factory(stepsize=istep, nused=nused, maxdist=maxdist, seed=77)
    pv.set_scene(scene)
    st = pv.Preprocess(wanted, stepsize=sstep, max_photon_depth=5, build=True)
    pos, wi, alpha, ids = pv.get_photons()
    ref = O.shoot(scene, wanted, sstep, istep, seed=77, rng_mode=O.PHILOX, nthreads=8)
    assert ref["rc"] == 0 and st.paths == ref["nshot"] and st.stack_overflows == 0
    common, ia, ib = np.intersect1d(ids, ref["ids"], return_indices=True)
    assert len(common) >= 0.995 * max(len(ids), len(ref["ids"]))
    dpos = np.abs(pos[ia] - ref["pos"][ib]).max(axis=1)
    assert np.quantile(dpos, 0.99) < 1e-4
    ok = dpos < 1e-4
    assert relerr(alpha[ia][ok], ref["alpha"][ib][ok]).max() < 1e-3
    # some of the photons do come from the area light: shooting with its slot dark gives another list
    dark = O.shoot(pkg.sceneio.read_scene(os.path.join(GOLDEN, "cornell_area.scn")), wanted, sstep, istep, seed=77, rng_mode=O.PHILOX, nthreads=8)
    assert dark["nshot"] != ref["nshot"] or len(dark["ids"]) != len(ref["ids"]) or not np.array_equal(dark["ids"], ref["ids"])
    # PhotonVolumeIntegrator::Li on the photons just shot: the direct term samples the area light
    rays = g["li_rays"]
    L, T = pv.Li(rays, ray_index_base=3)
    tree = O.KdTree(pos)
    oL, oT, ost = O.gather(scene, tree, wi, alpha, rays, istep, nused, maxdist, seed=77, ray_index_base=3)
    lit = oL > 1e-6 * oL.max()
    assert lit.any() and ost.shadow_rays > 0
    assert relerr(L, oL)[lit].max() < 1e-4 and np.abs(T - oT).max() < 1e-5


def test_area_light_scene_validation(pkg, pv_factory):
    import copy
    scene = area_scene(pkg, "volint_area")
    pv = pv_factory()
    s = copy.copy(scene)
    s.light_tris = None
    with pytest.raises(pkg.PVError) as e:
        pv.set_scene(s)
    assert "area light" in str(e.value)
    pv.set_scene(scene)
