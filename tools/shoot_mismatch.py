"""How close is the device shooter to the oracle on the same Philox streams, scene by scene?  (sets the bars of
tests/test_gpu_parity.py::test_shooter_vs_oracle_same_philox_stream from measurement instead of from caution)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from __graft_entry__ import load_package
pkg = load_package()
import oracle_lib as O
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, wanted, sstep in [("cornell_homog", 3000, 0.05), ("cornell_grid32", 1200, 0.05), ("rainbow_vol", 1500, 0.1), ("prism_small", 4000, 0.1),
                            ("sphere_glass", 4000, 0.1), ("sphere_disp", 4000, 0.1), ("cornell_exp", 1000, 0.05), ("underflow_glass", 3000, None)]:
    g = dict(np.load(os.path.join(G, name + ".npz"))); scene = pkg.sceneio.read_scene(os.path.join(G, name + ".scn"))
    istep = float(g["params"][2]); sstep = sstep or float(g["params"][1])
    pv = pkg.PhotonVolume(device=0, stepsize=istep, seed=77)
    pv.set_scene(scene)
    st = pv.Preprocess(wanted, stepsize=sstep, max_photon_depth=5, build=False)
    pos, wi, alpha, ids = pv.get_photons()
    ref = O.shoot(scene, wanted, sstep, istep, seed=77, rng_mode=O.PHILOX, nthreads=8)
    common, ia, ib = np.intersect1d(ids, ref["ids"], return_indices=True)
    dpos = np.abs(pos[ia] - ref["pos"][ib]).max(axis=1)
    rel = (np.abs(alpha[ia] - ref["alpha"][ib]) / np.maximum(np.abs(ref["alpha"][ib]), 1e-30)).max(axis=1)
    print("%-16s gpu %6d oracle %6d common %6d (only gpu %d, only oracle %d) nshot %s | pos max %.2e q99 %.2e | alpha relerr max %.2e q99 %.2e" % (
        name, len(ids), len(ref["ids"]), len(common), len(ids) - len(common), len(ref["ids"]) - len(common), st.paths == ref["nshot"],
        dpos.max(), np.quantile(dpos, 0.99), rel[dpos < 1e-4].max(), np.quantile(rel, 0.99)), flush=True)
    pv.close()
