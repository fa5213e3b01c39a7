"""Probe: shooter device time in the config-3 scene (tuning aid, not a bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from cs348b_pbrt_b200 import workloads as W
cfg = W.CONFIGS["config3"]
scene = W.load_scene(cfg)
pv = pkg.PhotonVolume(device=0, stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], seed=348)
pv.set_scene(scene)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t0 = time.perf_counter()
    st = pv.Preprocess(n, stepsize=0.05, max_photon_depth=5, build=False)
    dt = time.perf_counter() - t0
    print("%s: photons %d paths %d device %.4f s wall %.4f s -> %.1f M paths/s %.2f M photons/s" % (
        os.environ.get("PV_LIBPV", "default").split("/")[-1], pv.photon_count(), st.paths, st.seconds, dt, st.paths_local / st.seconds / 1e6, pv.photon_count() / st.seconds / 1e6), flush=True)
