"""A/B of shooter builds: same photons bit for bit?  device time per build.  usage: shoot_ab.py libA.so libB.so ... (first = baseline)"""
import os, sys, subprocess, json, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    pkg = load_package()
    from cs348b_pbrt_b200 import workloads as W
    cfg = W.CONFIGS["config3"]; scene = W.load_scene(cfg)
    pv = pkg.PhotonVolume(device=0, stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], seed=348)
    pv.set_scene(scene)
    best = 1e9
    for it in range(3):
        st = pv.Preprocess(int(sys.argv[3]), stepsize=0.05, max_photon_depth=5, build=False)
        best = min(best, st.seconds)
    pos, wi, alpha, ids = pv.get_photons()
    np.savez(sys.argv[2], pos=pos, wi=wi, alpha=alpha, ids=ids)
    ms = pv.PreprocessMaps(100000, 25000, 50000, True, stepsize=0.05)
    print(json.dumps({"device_s": best, "paths": int(st.paths), "photons": len(ids), "Mpaths_per_s": st.paths_local / best / 1e6,
                      "maps_device_s": ms.shoot.seconds, "maps_Mpaths_per_s": ms.shoot.paths_local / ms.shoot.seconds / 1e6}))
    sys.exit(0)
n = os.environ.get("SHOOT_N", "400000")
ref = None
for lib in sys.argv[1:]:
    out = "/tmp/ab_%s.npz" % os.path.basename(lib)
    env = dict(os.environ, PV_LIBPV=os.path.abspath(lib))
    r = subprocess.run([sys.executable, __file__, "--child", out, n], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    z = np.load(out)
    same = None
    if ref is None: ref = z
    else: same = all(np.array_equal(ref[k], z[k]) for k in ("pos", "wi", "alpha", "ids"))
    print(os.path.basename(lib), line[0] if line else r.stderr[-500:], "identical_to_first:", same, flush=True)
