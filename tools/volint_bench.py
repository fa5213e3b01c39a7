#!/usr/bin/env python
"""Throughput of pv_volume_li (VolumeIntegrator "single" / "emission", csrc/pv_volint.cu) on the scene shape of BASELINE
config 3: 256^3 density grid (made emitting), 1920x1080 camera rays in 8x8-tile order, stepsize 2/64.

Prints one JSON line per integrator: rays/s end to end through the C ABI with host buffers (copies inside), the device times
of the march kernels and of the recurrence kernel (CUDA events on the context's stream), the algorithmic bytes of each
(SURVEY.md 8(d) accounting: 32 B per density sample, 32 B per step record written / read, 272 B per ray) against the measured
HBM peak, and the pinned CPU oracle (one thread) on every 100th ray of the same frame.

    python tools/volint_bench.py [--steps 5] [--warmup 3] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def run_reference(args, cfg, scene, rays):
    """The reference's own classes and pthread task system over a sample of the frame (CPU only; needs oracle/_ref/ref_harness)."""
    import re
    import subprocess
    import tempfile
    from cs348b_pbrt_b200 import sceneio, scenes
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    sub = np.ascontiguousarray(rays[::args.ref_stride])
    threads = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        dens = os.path.join(tmp, "density.raw"); scene.density.astype(np.float32).tofile(dens)
        rf = os.path.join(tmp, "rays.bin"); sceneio.write_rays(rf, sub)
        for kind in ("single", "emission"):
            # the same medium as load_scene()'s (sigma_a 1, sigma_s 2, g 0.3) made emitting like the GPU arm's (Le 0.3 per bin would need
            # a spectrum; the grey "color Le" is close and the emission term costs the same)
            vol = scenes.grid_volume_text(32, scenes.blob_density(32)).replace('"float g"', '"color Le" [.3 .3 .3] "float g"')
            pbrt = os.path.join(tmp, kind + ".pbrt")
            open(pbrt, "w").write(scenes.volint_pbrt(kind, vol, stepsize=cfg["stepsize"]))
            ops = ["--ncores", str(threads), "--grid-file", str(cfg["grid"]), dens]
            for it in range(args.warmup + args.steps):
                ops += ["--li-parallel", rf, str(1000 + it), "-"]
            out = subprocess.run([harness, pbrt] + ops, capture_output=True, text=True)
            if out.returncode != 0:
                raise RuntimeError("ref_harness failed: " + out.stderr[-2000:])
            found = re.findall(r"li-parallel: (\d+) rays in ([0-9.]+) s on (\d+) cores", out.stderr)
            times = [float(t) for _, t, _ in found][args.warmup:]
            val = len(sub) * len(times) / sum(times)
            print(json.dumps({"impl": "reference", "metric": "volume-integrator rays/s", "integrator": kind, "value": val, "unit": "rays/s",
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
                              "config": {"workload": cfg["label"] + " (scene shape only)", "xres": cfg["xres"], "yres": cfg["yres"], "grid": cfg["grid"],
                                         "stepsize": cfg["stepsize"]},
                              "cpu_baseline": {"value": val, "unit": "rays/s", "cores": int(found[-1][2]), "kind": "reference",
                                               "sample": "every %d-th camera ray of the frame (%d rays) per step" % (args.ref_stride, len(sub))}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"],
                    help="reference: the UNMODIFIED reference's SingleScatteringIntegrator / EmissionIntegrator on all host cores "
                         "(oracle/_ref/ref_harness --li-parallel), same scene, every --ref-stride-th ray of the frame")
    ap.add_argument("--ref-stride", type=int, default=5)
    args = ap.parse_args()
    pkg = load_package()
    from cs348b_pbrt_b200 import workloads as W
    cfg = W.CONFIGS[args.workload]
    scene = W.load_scene(cfg)
    for b in range(pkg._abi.NSPEC):
        scene.medium.le[b] = 0.3
    rays, _ = W.frame_rays(cfg)
    n = len(rays)
    if args.impl == "reference":
        return run_reference(args, cfg, scene, rays)
    peak = 6454.9
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    pv = pkg.PhotonVolume(device=0, stepsize=cfg["stepsize"], seed=1)
    pv.set_scene(scene)
    for kind in ("single", "emission"):
        for _ in range(args.warmup):
            L, T = pv.VolumeLi(kind, rays)
        pv.gather_stats(reset=True)
        wall, kern, march = [], [], []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            L, T = pv.VolumeLi(kind, rays)
            wall.append(time.perf_counter() - t0)
            kern.append(pv.last_kernel_ms()); march.append(pv.last_march_ms())
        st = pv.gather_stats(reset=True)
        # march steps of the frame, from the records the kernels agree on: every step with non-zero density costs one density
        # sample at the sample point; count them exactly on the host instead (medium box [-1,1]^3)
        o = rays["o"].astype(np.float64); d = rays["d"].astype(np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            ta = (-1.0 - o) / d; tb = (1.0 - o) / d
        t0r = np.maximum(np.minimum(ta, tb).max(axis=1), rays["mint"]); t1r = np.minimum(np.maximum(ta, tb).min(axis=1), rays["maxt"])
        hit = t1r > t0r
        steps = int(np.ceil((t1r[hit] - t0r[hit]) / cfg["stepsize"]).sum())
        k_ms, m_ms, w_s = float(np.mean(kern)), float(np.mean(march)), float(np.mean(wall))
        dens = st.density_samples / args.steps
        march_bytes = dens * 32 + steps * 32 + n * (40 + 32)
        recur_bytes = steps * 32 + n * (32 + 24 + 240)
        line = {
            "metric": "volume-integrator rays/s", "integrator": kind, "unit": "rays/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "config": {"workload": cfg["label"] + " (scene shape only: no photon map; medium made emitting)", "xres": cfg["xres"],
                       "yres": cfg["yres"], "grid": cfg["grid"], "stepsize": cfg["stepsize"], "ray_order": "8x8 tiles"},
            "e2e": {"value": n / w_s, "unit": "rays/s", "h2d_bytes_per_step": int(rays.nbytes), "d2h_bytes_per_step": int(L.nbytes + T.nbytes)},
            "value": n / ((k_ms + m_ms) * 1e-3), "ms_per_step": k_ms + m_ms,
            "march_steps": steps, "density_samples_per_pass": dens, "shadow_rays_per_pass": st.shadow_rays / args.steps,
            "kernels": {
                "march (march_setup + march_steps)": {"ms": m_ms, "algorithmic_bytes": march_bytes, "achieved_gbs": march_bytes / (m_ms * 1e-3) / 1e9,
                                                       "frac_of_hbm_peak": march_bytes / (m_ms * 1e-3) / 1e9 / peak},
                "volint_kernel": {"ms": k_ms, "algorithmic_bytes": recur_bytes, "achieved_gbs": recur_bytes / (k_ms * 1e-3) / 1e9,
                                  "frac_of_hbm_peak": recur_bytes / (k_ms * 1e-3) / 1e9 / peak},
            },
            "hbm_peak_gbs": peak, "checksum_L": float(L.sum(dtype=np.float64)), "mean_T": float(T.mean()),
        }
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            sub = np.ascontiguousarray(rays[::100])
            t0 = time.perf_counter()
            oL, oT, _ = O.volume_li(scene, sub, cfg["stepsize"], O.SINGLE if kind == "single" else O.EMISSION, seed=1)
            dt = time.perf_counter() - t0
            # same stream only if the ray indices match: the sample is every 100th ray, so compare through a second device call
            gL, gT = pv.VolumeLi(kind, sub)
            m = oL > 0
            line["cpu_baseline"] = {"value": len(sub) / dt, "unit": "rays/s", "cores": 1, "kind": "port",
                                    "sample": "every 100th camera ray (%d rays), %.1f s" % (len(sub), dt),
                                    "max_rel_err_vs_gpu": float((np.abs(gL - oL)[m] / oL[m]).max()) if m.any() else 0.0}
        print(json.dumps(line), flush=True)
    pv.close()


if __name__ == "__main__":
    main()
