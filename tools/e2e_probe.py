"""Probe: where does the host-pointer pv_gather spend its time (tuning aid, not a bench)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
import torch
from cs348b_pbrt_b200 import workloads as W
cfg = W.CONFIGS["config3"]
scene = W.load_scene(cfg)
cache = "/tmp/pvcache/ph_config3_16000000_0_16000000.npz"
if os.path.exists(cache):
    z = np.load(cache); pos, wi, alpha = z["pos"], z["wi"], z["alpha"]
else:
    pos, wi, alpha = W.photons_from_density(scene, cfg["photons"])
pv = pkg.PhotonVolume(device=0, stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], seed=348)
pv.set_scene(scene); pv.set_photons(pos, wi, alpha); pv.build()
rays, _ = W.frame_rays(cfg)
n = len(rays)
h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 10).copy()).pin_memory()
h_L = torch.empty((n, 30)).pin_memory(); h_T = torch.empty((n, 30)).pin_memory()
d_rays = h_rays.cuda(); d_L = torch.empty((n, 30), device="cuda"); d_T = torch.empty((n, 30), device="cuda")
for s in ["dev", "1", "2", "4", "8"]:
    if s != "dev":
        os.environ["PV_GATHER_SLICES"] = s
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if s == "dev":
            pv.Li_dev(d_rays, n, d_L, d_T)
        else:
            pv.Li_into(h_rays, n, h_L, h_T)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("slices=%s wall %.2f ms  gather %.2f ms  march %.2f ms" % (s, dt * 1e3, pv.last_kernel_ms(), pv.last_march_ms()), flush=True)
# raw copy speed
torch.cuda.synchronize(); t0 = time.perf_counter(); d_L.copy_(h_L, non_blocking=True); torch.cuda.synchronize(); print("H2D 249MB %.2f ms" % ((time.perf_counter() - t0) * 1e3))
torch.cuda.synchronize(); t0 = time.perf_counter(); h_L.copy_(d_L, non_blocking=True); torch.cuda.synchronize(); print("D2H 249MB %.2f ms" % ((time.perf_counter() - t0) * 1e3))
