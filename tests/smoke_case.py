"""The case __graft_entry__.smoke() runs: one small gather on cuda:0, checked against the CPU oracle."""
import os
import numpy as np


def run(pkg):
    import oracle_lib as O
    here = os.path.dirname(os.path.abspath(__file__))
    g = dict(np.load(os.path.join(here, "golden", "cornell_homog.npz")))
    scene = pkg.sceneio.read_scene(os.path.join(here, "golden", "cornell_homog.scn"))
    stepsize, nused, maxdist = float(g["params"][2]), int(g["params"][0]), float(g["params"][1])
    pv = pkg.PhotonVolume(device=0, stepsize=stepsize, nused=nused, maxdist=maxdist, seed=1)
    pv.set_scene(scene)
    # shoot a few photons on the GPU, build the map, gather 96 rays
    st = pv.Preprocess(2000, stepsize=0.05)
    pos, wi, alpha, ids = pv.get_photons()
    rays = g["li_rays"]
    L, T = pv.Li(rays)
    tree = O.KdTree(pos)
    oL, oT, _ = O.gather(scene, tree, wi, alpha, rays, stepsize, nused, maxdist, seed=1)
    m = oL > 0
    err = float((np.abs(L - oL)[m] / oL[m]).max()) if m.any() else 0.0
    assert m.any() and err < 1e-4, "smoke: radiance mismatch %g" % err
    assert np.abs(T - oT).max() < 1e-5
    print("smoke ok: %d photons from %d paths, %d rays, max rel err %.2e" % (len(pos), st.paths, len(rays), err))
    pv.close()
