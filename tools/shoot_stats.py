"""Probe: per-path work counters of the shooter in the config-3 scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from cs348b_pbrt_b200 import workloads as W
cfg = W.CONFIGS["config3"]; scene = W.load_scene(cfg)
pv = pkg.PhotonVolume(device=0, stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], seed=348)
pv.set_scene(scene)
st = pv.Preprocess(int(sys.argv[1]) if len(sys.argv) > 1 else 400000, stepsize=0.05, max_photon_depth=5, build=False)
p = st.paths_local
print("paths %d segments/path %.2f density samples/path %.1f nodes/path %.1f tris/path %.1f device %.4f s" % (
    p, st.segments / p, st.density_samples / p, st.nodes_visited / p, st.tri_tests / p, st.seconds))
print("integrator stepsize", cfg["stepsize"], "-> tau step", 4 * cfg["stepsize"])
