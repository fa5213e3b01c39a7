"""Probe: pv_build_bvh on a synthetic mesh (n triangles, half of them crowded into 1/1000 of the volume) -- build time per call and
traversal cost through the tree (nodes / triangle tests per ray, from pv_intersect's wall time).  usage: lbvh_probe.py [n] [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from __graft_entry__ import load_package
pkg = load_package()
from test_gpu_lbvh import soup
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tri = soup(n, 11, size=0.004)
v = tri.reshape(-1, 3, 3)
bounds = np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1).astype(np.float32)
pv = pkg.PhotonVolume(device=0)
for it in range(reps):
    t0 = time.perf_counter()
    nodes, order, ms = pv.build_bvh(bounds, 4)
    print("pv_build_bvh: %d primitives -> %d nodes, kernels %.3f ms, call %.1f ms (%.0f B of algorithmic traffic per primitive -> %.0f GB/s)" % (
        n, len(nodes) // 32, ms, 1e3 * (time.perf_counter() - t0), 512, 512.0 * n / (ms * 1e-3) / 1e9), flush=True)
