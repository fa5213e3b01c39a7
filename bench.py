#!/usr/bin/env python
"""bench.py -- volumetric photon-mapping hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

Workload (config.workload): BASELINE.json configs[2] -- Cornell box + heterogeneous 256^3 density-grid medium,
16M photons, 1920x1080 camera rays, fixed-radius gather (nused 512 >= photons in radius, maxdist 0.018,
stepsize 2/64).  Metric: volume-gather rays/s (whole job, all GPUs); photons-traced/s of the shooter is
reported in the same line under "shoot".  A step = one gather pass (PhotonVolumeIntegrator::Li for every camera
ray of the frame) against the resident photon map.  N > 1: photons are produced sharded, the map is replicated
with one NCCL all-gather over NVLink, image tiles are dealt round-robin to ranks (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []; self.proc = None; self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_gather_baseline(W, cfg, scene, pos, wi, alpha, rays, nsample, threads, seed):
    """The CPU port (oracle) timed on a bounded sample of the same rays; kd-tree build is untimed preprocess."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    t0 = time.perf_counter()
    tree = O.KdTree(pos)
    build_s = time.perf_counter() - t0
    sel = np.linspace(0, len(rays) - 1, nsample).astype(np.int64)
    sample = np.ascontiguousarray(rays[sel])
    t0 = time.perf_counter()
    L, T, st = O.gather(scene, tree, wi, alpha, sample, cfg["stepsize"], cfg["nused"], cfg["maxdist"], seed=seed, nthreads=threads)
    dt = time.perf_counter() - t0
    return {"value": nsample / dt, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": "%d of %d camera rays (every %d-th), full %d-photon map; kd-tree build %.1f s untimed; %.1f s timed"
                      % (nsample, len(rays), max(1, len(rays) // nsample), len(pos), build_s, dt),
            "lookups_per_s": st.lookups / dt}, (sel, L, T)


def cpu_shoot_baseline(scene, cfg, n_wanted, threads, seed):
    """The CPU port's photon shooter (oracle; per-path Philox streams, 4096-path blocks spread over the host threads) on a bounded
    sample in the same scene with the same parameters as the GPU sample above: the `shoot` rates' CPU counterpart."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    O.shoot(scene, 2000, 0.05, cfg["stepsize"], 5, seed=seed, rng_mode=O.PHILOX, nthreads=threads)       # untimed: library load, thread start
    t0 = time.perf_counter()
    r = O.shoot(scene, n_wanted, 0.05, cfg["stepsize"], 5, seed=seed, rng_mode=O.PHILOX, nthreads=threads)
    dt = time.perf_counter() - t0
    return {"paths_per_s": r["nshot"] / dt, "photons_per_s": r["n"] / dt, "cores": threads, "kind": "port",
            "sample": "%d volume photons from %d light paths, %.1f s" % (r["n"], r["nshot"], dt)}


def run_reference(args, cfg, W, scene):
    """--impl reference: the reference's CPU implementation of the gather on the host cores (rank 0 only).
    Preferred: oracle/_ref/ref_harness = the UNMODIFIED reference (KdTree<Photon>, PhotonVolumeIntegrator::Li, its pthread
    task system) built from /root/reference by oracle/Makefile; it is fed the same scene, photon set and camera rays through
    files.  Fallback when that binary is absent: the C port (oracle/pv_oracle.c)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import re
    import tempfile
    threads = os.cpu_count() or 1
    n_ph = args.photons or cfg["photons"]
    pos, wi, alpha = W.photons_from_density(scene, n_ph)
    rays, _ = W.frame_rays(cfg)
    nsample = args.ref_rays
    nit = args.warmup + args.steps
    samples = []
    for it in range(nit):
        sel = (np.linspace(0, len(rays) - 1, nsample).astype(np.int64) + it * 7) % len(rays)
        samples.append(np.ascontiguousarray(rays[sel]))
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    kind = "reference" if os.path.exists(harness) and cfg["grid"] else "port"
    times = []
    shoot_ref = None
    if kind == "reference":
        from cs348b_pbrt_b200 import sceneio, scenes
        with tempfile.TemporaryDirectory() as tmp:
            pho = os.path.join(tmp, "photons.pho")
            with open(pho, "wb") as f:
                f.write(b"PVPHOT01" + np.uint64(n_ph).tobytes())
                for a in range(0, n_ph, 1 << 20):
                    b = min(n_ph, a + (1 << 20))
                    f.write(np.concatenate([pos[a:b], wi[a:b], alpha[a:b]], axis=1).astype(np.float32).tobytes())
            dens = os.path.join(tmp, "density.raw")
            scene.density.astype(np.float32).tofile(dens)
            pbrt = os.path.join(tmp, "scene.pbrt")
            open(pbrt, "w").write(scenes.cornell_pbrt(scenes.grid_volume_text(32, scenes.blob_density(32)), 0, stepsize=cfg["stepsize"],
                                                      nused=cfg["nused"], maxdist=cfg["maxdist"]))
            ops = ["--ncores", str(threads), "--grid-file", str(cfg["grid"]), dens, "--load-photons", pho]
            for it in range(nit):
                rf = os.path.join(tmp, "rays_%d.bin" % it)
                sceneio.write_rays(rf, samples[it])
                ops += ["--li-parallel", rf, str(1000 + it), "-"]
            out = subprocess.run([harness, pbrt] + ops, capture_output=True, text=True)
            if out.returncode != 0:
                raise RuntimeError("ref_harness failed: " + out.stderr[-2000:])
            found = re.findall(r"li-parallel: (\d+) rays in ([0-9.]+) s on (\d+) cores", out.stderr)
            times = [float(t) for _, t, _ in found][args.warmup:]
            if found:
                threads = int(found[-1][2])
            # the reference's own multi-threaded photon shooting (PhotonShootingTask x cores, core/photonshooter.cpp:232-357) in the same
            # scene with the parameters of the GPU arm's shooting sample: the `shoot` rates' reference counterpart
            if args.cpu_shoot_photons > 0:
                pbrt2 = os.path.join(tmp, "shoot.pbrt")
                open(pbrt2, "w").write(scenes.cornell_pbrt(scenes.grid_volume_text(32, scenes.blob_density(32)), args.cpu_shoot_photons,
                                                           stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], shoot_step=0.05))
                o2 = subprocess.run([harness, pbrt2, "--ncores", str(threads), "--grid-file", str(cfg["grid"]), dens, "--shoot"],
                                    capture_output=True, text=True)
                m = re.search(r"shot: nshot=(\d+) volume=(\d+) .*tasks=(\d+) seconds=([0-9.]+)", o2.stderr)
                if o2.returncode == 0 and m and float(m.group(4)) > 0:
                    nshot, nvol, ntask, sec = int(m.group(1)), int(m.group(2)), int(m.group(3)), float(m.group(4))
                    shoot_ref = {"paths_per_s": nshot / sec, "photons_per_s": nvol / sec, "cores": ntask, "kind": "reference",
                                 "sample": "%d volume photons from %d light paths, %.1f s (shooter stepsize 0.05, maxphotondepth 5)" % (nvol, nshot, sec)}
    if not times:
        kind = "port"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        tree = O.KdTree(pos)
        for it in range(nit):
            t0 = time.perf_counter()
            O.gather(scene, tree, wi, alpha, samples[it], cfg["stepsize"], cfg["nused"], cfg["maxdist"], seed=args.seed, nthreads=threads)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    dt = sum(times)
    val = nsample * len(times) / dt
    out = {"impl": "reference", "metric": "volume-gather rays/s", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / len(times), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": cfg["label"], "photons": n_ph, "grid": cfg["grid"], "xres": cfg["xres"], "yres": cfg["yres"],
                      "stepsize": cfg["stepsize"], "nused": cfg["nused"], "maxdist": cfg["maxdist"]},
           "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": kind,
                            "sample": "%d camera rays per step (every %d-th of the %d-ray frame), full %d-photon map in the reference's KdTree"
                                      % (nsample, max(1, len(rays) // nsample), len(rays), n_ph)},
           "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "shoot": shoot_ref}
    print(json.dumps(out), flush=True)


def extra_line(pkg, W, name, local, args, peak):
    """A secondary workload through the same path, device-resident, one GPU: BASELINE config 2 (Cornell box + homogeneous medium,
    1 M photons, 512x512, k = 50 nearest within 0.25 -- the k-nearest regime the reference's shipped scenes use)."""
    import torch
    cfg = W.CONFIGS[name]
    scene = W.load_scene(cfg)
    pv = pkg.PhotonVolume(device=local, stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], seed=args.seed)
    try:
        pv.set_scene(scene)
        pos, wi, alpha = W.photons_from_density(scene, cfg["photons"])
        pv.set_photons(pos, wi, alpha); pv.build()
        rays, _ = W.frame_rays(cfg)
        n = len(rays)
        dev = torch.device("cuda", local)
        d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 10)).to(dev)
        d_L = torch.empty((n, 30), device=dev); d_T = torch.empty((n, 30), device=dev)
        for _ in range(3):
            pv.Li_dev(d_rays, n, d_L, d_T)
        pv.gather_stats(reset=True)
        ext = torch.cuda.ExternalStream(pv.stream(), device=dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        steps = 5; kms = []
        torch.cuda.synchronize()
        e0.record(ext)
        for _ in range(steps):
            pv.Li_dev(d_rays, n, d_L, d_T); kms.append(pv.last_kernel_ms())
        e1.record(ext)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        st = pv.gather_stats(reset=True)
        byts = W.gather_bytes(st, n * steps) / steps
        ach = byts / (float(np.mean(kms)) * 1e-3) / 1e9
        return {"workload": cfg["label"], "value": n / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "steps": steps,
                "lookups_per_step": st.lookups / steps, "photons_found_per_lookup": st.photons_found / max(st.lookups, 1),
                "candidates_per_lookup": st.candidates_tested / max(st.lookups, 1),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "avg_launch_ms": float(np.mean(kms)), "algorithmic_bytes_per_launch": byts},
                "L_sum": float(d_L.sum().item())}
    finally:
        pv.close()


def lbvh_line(pkg, local, peak, n=2_000_000):
    """pv_build_bvh (csrc/pv_lbvh.cu) on a synthetic mesh of n small triangles, half of them crowded into 1/1000 of the
    volume: device time of the build's kernels (CUDA events inside the library), second call (warm), and the call's wall time with
    its host copies.  Algorithmic bytes: DESIGN.md 4 K6 (~512 B per primitive)."""
    import time
    rs = np.random.RandomState(11)
    c = rs.uniform(-1, 1, (n, 1, 3)).astype(np.float32)
    c[: n // 2] = c[: n // 2] * np.float32(0.1) + np.float32(0.4)
    v = c + rs.uniform(-0.004, 0.004, (n, 3, 3)).astype(np.float32)
    bounds = np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1).astype(np.float32)
    pv = pkg.PhotonVolume(device=local)
    try:
        pv.build_bvh(bounds, 4)
        t0 = time.perf_counter()
        nodes, order, ms = pv.build_bvh(bounds, 4)
        wall = time.perf_counter() - t0
        ach = 512.0 * n / (ms * 1e-3) / 1e9
        return {"workload": "LBVH over %d synthetic triangles, leaves of <= 4" % n, "primitives": n, "nodes": len(nodes) // 32, "kernels_ms": ms,
                "call_wall_ms": wall * 1e3, "primitives_per_s": n / (ms * 1e-3),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "algorithmic_bytes_per_launch": 512.0 * n, "note": "27 launches; launch / latency bound at this size"}}
    finally:
        pv.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--photons", type=int, default=0, help="override the photon count (debug only; invalidates the number)")
    ap.add_argument("--shoot-photons", type=int, default=-1, help="photons of the full-frame pass (shoot -> all-gather -> build -> gather); -1 = the workload's photon count, 0 = skip")
    ap.add_argument("--maps-photons", type=int, default=200_000, help="volume-photon target of the all-maps shooting sample (0 = skip)")
    ap.add_argument("--cpu-rays", type=int, default=600_000, help="rays of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-shoot-photons", type=int, default=150_000, help="photons of the bounded CPU-baseline shooting sample (0 = skip)")
    ap.add_argument("--ref-rays", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workload lines (config 2: k-nearest gather)")
    ap.add_argument("--seed", type=int, default=348)
    ap.add_argument("--slice-of", type=int, default=0, help="tuning aid: gather only the rays rank 0 of that many ranks would get (one process; invalidates the number)")
    args = ap.parse_args()

    pkg = load_package()
    from cs348b_pbrt_b200 import workloads as W, multigpu as MG
    cfg = W.CONFIGS[args.workload]
    scene = W.load_scene(cfg)
    if args.impl == "reference":
        run_reference(args, cfg, W, scene)
        return

    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    A = pkg._abi
    peak, peak_kind = read_peaks()
    n_ph = args.photons or cfg["photons"]

    pv = pkg.PhotonVolume(device=local, stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], seed=args.seed)
    pv.set_scene(scene)
    ext = torch.cuda.ExternalStream(pv.stream(), device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    import ctypes as C

    def shoot_pass(target, wave0):
        """PhotonShooter::Preprocess, volume branch, sharded by 4096-path block over the ranks (pv_shoot_blocks / pv_shoot_finish):
        returns (stats, wall seconds).  Per-wave deposit counts are summed over ranks so every rank stops at the same block."""
        prm = A.ShootParams(0.05, cfg["stepsize"], 5, args.seed, rank, world, 0, 0.0)
        st = A.ShootStats()
        block, total, wave, last = 0, 0, wave0, 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        while not last:
            counts = (C.c_uint32 * wave)()
            pv._chk(pv.lib.pv_shoot_blocks(pv.ctx, C.c_uint64(block + 1), C.c_uint32(wave), C.byref(prm), counts, C.byref(st)))
            cnt = np.ctypeslib.as_array(counts).astype(np.int64)
            if dist is not None:
                t = torch.from_numpy(cnt).to(dev); dist.all_reduce(t); cnt = t.cpu().numpy()
            last, total, used = MG.last_block(cnt, block + 1, total, target)
            block += used
            if not last:
                wave = MG.next_wave(total, block, target, world)
        pv._chk(pv.lib.pv_shoot_finish(pv.ctx, C.c_uint64(last)))
        torch.cuda.synchronize()
        return st, last, time.perf_counter() - t0

    def shoot_report(st, last, wall, target):
        sec = torch.tensor([st.seconds, float(st.paths_local), float(pv.photon_count()), float(st.nodes_visited), float(st.tri_tests),
                            float(st.density_samples), float(st.stack_overflows)], dtype=torch.float64, device=dev)
        mx = sec.clone()
        if dist is not None:
            dist.all_reduce(sec); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sec = sec.cpu().numpy(); dsec = float(mx.cpu().numpy()[0])
        shoot_bytes = sec[3] * 32 + sec[4] * 36 + sec[5] * 32 + sec[2] * 144       # SURVEY 8d shoot formula
        ach = shoot_bytes / dsec / 1e9
        return {"photons": int(sec[2]), "paths": int(last * 4096), "paths_traced_incl_discarded": int(sec[1]), "device_s": dsec, "wall_s": wall,
                "paths_per_s": sec[1] / dsec, "photons_per_s": sec[2] / dsec, "stack_overflows": int(sec[6]),
                "roofline": {"bound": "hbm", "kernel": "wavefront shooter (wf_trace_kernel + wf_march_kernel + wf_event_kernel per generation)", "achieved": ach / world, "peak": peak, "unit": "GB/s", "frac": ach / (peak * world),
                             "peak_kind": peak_kind, "algorithmic_bytes": shoot_bytes, "traffic": None,
                             "note": "per GPU; bytes = nodes*32 + triangle tests*36 + density samples*32 + deposits*144 (SURVEY 8d), counted by the kernels "
                                     "for the work they do (transmittance marches whose result nothing reads are not done); the kernels are "
                                     "instruction-issue-bound on the trilinear sampler, not HBM-bound (profiles/r02_*_summary.md)"},
                "hbm_frac_algorithmic": ach / (peak * world),
                "params": {"shooter_stepsize": 0.05, "maxphotondepth": 5, "target": target}}

    # ------------------------------------------------------------------ rays of this rank (image tiles dealt round-robin)
    rays, order = W.frame_rays(cfg, rank, world, density=scene.density) if not args.slice_of else W.frame_rays(cfg, 0, args.slice_of, density=scene.density)
    n_local = len(rays)
    n_total = cfg["xres"] * cfg["yres"]
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 10)).to(dev)
    d_L = torch.empty((n_local, 30), device=dev); d_T = torch.empty((n_local, 30), device=dev)

    # ------------------------------------------------------------------ the whole frame once, at FULL size: shoot -> all-gather -> build -> gather
    # (the photon map the gather reads here is the one the shooter just made; NCCL and every kernel are warm before the timers start)
    shoot, frame = {}, None
    if dist is not None:
        warm = torch.zeros(1, device=dev); dist.all_reduce(warm); torch.cuda.synchronize()     # lazy NCCL initialisation happens here
        # the library's own communicator (csrc/pv_comm.cu): rank 0 makes the NCCL id, torch.distributed only carries its 128 bytes
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid = torch.tensor(list(pkg.PhotonVolume.comm_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        pv.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    if args.shoot_photons != 0:
        target = n_ph if args.shoot_photons < 0 else args.shoot_photons
        shoot_pass(min(target, 20000 * world), 8 * world)                                       # untimed: module load, local-memory reservation
        if dist is not None:
            pv.allgather_photons()                                                              # untimed: the communicator's first collective
        pv.build(); pv.Li_dev(d_rays, n_local, d_L, d_T)                                        # untimed: the gather's step / L_ii / sort buffers (about 10 GB of cudaMalloc for a 1080p frame,
                                                                                                # 3-190 ms depending on the driver's mood) exist before the frame is timed, as in a renderer's second frame
        barrier()
        t_frame = time.perf_counter()
        st, last, wall = shoot_pass(target, 64 * world)
        shoot = shoot_report(st, last, wall, target)
        n_shot_local = pv.photon_count()
        ag = None
        if dist is not None:
            # replicate the shot photons: pv_allgather_photons = one grouped ncclAllGather per SoA plane, straight from the shooter's
            # planes into the planes pv_build reads, then the sort by photon id that makes the set independent of the rank count
            t0 = time.perf_counter()
            ag_ms = pv.allgather_photons()
            torch.cuda.synchronize()
            n_all = pv.photon_count()
            nbytes = (world - 1) / world * n_all * 160
            ag = {"collective_ms": ag_ms, "wall_ms": (time.perf_counter() - t0) * 1e3, "photons": n_all, "bytes_in_per_gpu": nbytes,
                  "gb_per_s_per_gpu": nbytes / (ag_ms * 1e-3) / 1e9 if ag_ms > 0 else None}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pv.build()
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        pv.Li_dev(d_rays, n_local, d_L, d_T)
        torch.cuda.synchronize()
        gather_wall_ms = (time.perf_counter() - t0) * 1e3
        gather_dev_ms = pv.last_kernel_ms() + pv.last_march_ms()
        gather_phases = dict(zip(("step_sort", "cellgather_kernel", "overflow_pass", "recurrence"), pv.last_phase_ms())); gather_phases["march_kernels"] = pv.last_march_ms()
        barrier()
        frame_wall = time.perf_counter() - t_frame
        vals = torch.tensor([shoot["device_s"] * 1e3, wall * 1e3, build_ms, gather_dev_ms, gather_wall_ms, frame_wall * 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        v = [float(x) for x in vals.cpu().numpy()]
        frame = {"what": "ONE frame end to end at full size, max over ranks: shoot %d photons (sharded by 4096-path block) -> all-gather -> grid build -> "
                         "gather of the frame's camera rays against the map just shot" % shoot["photons"],
                 "shoot_device_ms": v[0], "shoot_wall_ms": v[1], "allgather": ag, "build_wall_ms": v[2], "gather_device_ms": v[3],
                 "gather_wall_ms": v[4], "gather_phase_ms_rank0": gather_phases, "frame_wall_ms": v[5], "gather_L_finite": bool(torch.isfinite(d_L).all().item()),
                 "gather_L_sum": float(d_L.sum().item())}

        # the same pass with the SURFACE maps on (pv_shoot_maps + pv_radiance_photons): every photon class of
        # the reference's shooter in the config-3 scene, bounded sample, single rank (rank 0 reports)
        if world == 1 and args.maps_photons > 0:
            pvm = pkg.PhotonVolume(device=local, stepsize=cfg["stepsize"], nused=cfg["nused"], maxdist=cfg["maxdist"], seed=args.seed)
            pvm.set_scene(scene)
            n = args.maps_photons
            pvm.PreprocessMaps(2000, 500, 1000, True, stepsize=0.05, max_photon_depth=5)      # untimed warm-up
            t0 = time.perf_counter()
            ms = pvm.PreprocessMaps(n, n // 4, n // 2, True, stepsize=0.05, max_photon_depth=5)
            t1 = time.perf_counter()
            Lo = pvm.RadiancePhotons(50, 0.05 ** 2)
            t2 = time.perf_counter()
            shoot["all_maps"] = {"photons": {k: int(ms.n[i]) for i, k in enumerate(("volume", "caustic", "indirect", "direct", "radiance"))},
                                 "paths": int(ms.nshot), "replayed_blocks": int(ms.replayed_blocks), "device_s": float(ms.shoot.seconds),
                                 "paths_per_s": float(ms.shoot.paths_local) / float(ms.shoot.seconds),
                                 "photons_per_s": float(sum(ms.n)) / float(ms.shoot.seconds), "wall_s": t1 - t0,
                                 "radiance_photons_wall_s": t2 - t1, "radiance_finite": bool(np.isfinite(Lo).all()),
                                 "params": {"volume": n, "caustic": n // 4, "indirect": n // 2, "finalgather": True, "nlookup": 50, "maxdist": 0.05}}
            pvm.close()

    # ------------------------------------------------------------------ the photon map of the timed gather workload (same synthetic set as the reference arm)
    lo, hi = W.photon_slice(n_ph, rank, world)
    t0 = time.perf_counter()
    cache = os.environ.get("PV_BENCH_CACHE")               # tuning runs: keep the generated photon slice between processes
    cfile = os.path.join(cache, "ph_%s_%d_%d_%d.npz" % (args.workload, n_ph, lo, hi)) if cache else None
    if cfile and os.path.exists(cfile):
        z = np.load(cfile); pos, wi, alpha = z["pos"], z["wi"], z["alpha"]
    else:
        pos, wi, alpha = W.photons_from_density(scene, n_ph, lo=lo, hi=hi)
        if cfile:
            os.makedirs(cache, exist_ok=True); np.savez(cfile, pos=pos, wi=wi, alpha=alpha)
    gen_s = time.perf_counter() - t0
    allgather = None
    pv.set_photons(pos, wi, alpha)                             # this rank's slice of the set
    if world > 1:
        barrier()
        ag_ms = pv.allgather_photons(renumber=True)            # union in global photon order, photon i keeps index i
        assert pv.photon_count() == n_ph
        t = torch.tensor([ag_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = (world - 1) / world * n_ph * 160
        allgather = {"ms": float(t.item()), "bytes_in_per_gpu": nbytes, "gb_per_s_per_gpu": nbytes / (float(t.item()) * 1e-3) / 1e9,
                     "nvlink_peer_gb_per_s": 770.0, "how": "pv_allgather_photons: grouped ncclAllGather of the four SoA planes (160 B per photon)"}
    torch.cuda.synchronize()
    pv.build()                                                 # first build of this set: sizes the map's buffers
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pv.build()                                                 # timed on a warm context, inputs resident, no allocation inside
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0

    for _ in range(max(args.warmup, 3)):
        pv.Li_dev(d_rays, n_local, d_L, d_T)
    pv.gather_stats(reset=True)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    kernel_ms = []; march_ms = []; phase_ms = []
    launches0 = pv.launch_count()
    e0.record(ext)
    for _ in range(args.steps):
        pv.Li_dev(d_rays, n_local, d_L, d_T)
        kernel_ms.append(pv.last_kernel_ms()); march_ms.append(pv.last_march_ms()); phase_ms.append(pv.last_phase_ms())
    e1.record(ext)
    launches = pv.launch_count() - launches0
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    stats = pv.gather_stats(reset=True)
    value = n_total * args.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel, this rank: cellgather_kernel (cell-batched schedule: lookups + flux sums of every march step)
    # or gather_kernel (k-nearest regime: lookups fused with the recurrence), timed with CUDA events on the library's stream
    bytes_per_launch = W.gather_bytes(stats, n_local * args.steps) / args.steps
    ph = dict(zip(("step_sort", "cellgather_kernel", "overflow_pass", "recurrence"), [float(v) for v in np.mean(phase_ms, axis=0)]))
    cell = ph["cellgather_kernel"] > 0
    avg_kernel_ms = ph["cellgather_kernel"] if cell else float(np.mean(kernel_ms))
    achieved = bytes_per_launch / (avg_kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_kind": peak_kind, "kernel": "cellgather_kernel" if cell else "gather_kernel", "avg_launch_ms": avg_kernel_ms,
                "gather_ms": float(np.mean(kernel_ms)), "march_kernels_ms": float(np.mean(march_ms)), "phase_ms": ph,
                "algorithmic_bytes_per_launch": bytes_per_launch, "b_ph": 144,
                "lookups_per_launch": stats.lookups / args.steps, "photons_found_per_lookup": stats.photons_found / max(stats.lookups, 1),
                "candidates_per_lookup": stats.candidates_tested / max(stats.lookups, 1),
                "candidates_staged_per_lookup": stats.candidates_tested / max(stats.lookups, 1) / (32.0 if cell else 1.0),
                "note": "achieved = SURVEY 8d algorithmic bytes (every photon record a lookup uses counts, sum nFound*144 + lookups*28 + rays*272) / "
                        "the kernel's CUDA-event time.  The cell-batched kernel stages a block of cells ONCE for 32 lookups and its alpha lines hit "
                        "L1/L2, so the figure can exceed the HBM peak; traffic (dram bytes per launch) is not measurable inside this run: see the "
                        "ncu captures under profiles/ (r02_*_summary.md)"}

    # ------------------------------------------------------------------ end to end through the host-pointer C ABI call
    h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 10).copy()).pin_memory()
    h_L = torch.empty((n_local, 30)).pin_memory(); h_T = torch.empty((n_local, 30)).pin_memory()
    pv.Li_into(h_rays, n_local, h_L, h_T)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pv.Li_into(h_rays, n_local, h_L, h_T)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = {"value": n_total * args.steps / float(e2e_s.item()), "unit": "rays/s", "h2d_bytes_per_step": int(n_local * 40),
           "d2h_bytes_per_step": int(n_local * 240)}
    check = float(h_L.sum().item())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu, (sel, cL, cT) = cpu_gather_baseline(W, cfg, scene, pos, wi, alpha, rays, args.cpu_rays, threads, args.seed)
        # same sample through the CUDA path with the same per-ray Philox indices (ray_index_base 0)
        gL, gT = pv.Li(np.ascontiguousarray(rays[sel]))
        m = cL > 0
        cpu["max_rel_err_vs_gpu"] = float((np.abs(gL - cL)[m] / cL[m]).max()) if m.any() else 0.0
        if shoot and args.cpu_shoot_photons > 0:
            try:
                shoot["cpu_baseline"] = cpu_shoot_baseline(scene, cfg, args.cpu_shoot_photons, threads, args.seed)
            except Exception as e:                          # a reported baseline must never cost the bench line
                shoot["cpu_baseline"] = {"error": repr(e)}

    extra = None
    if world == 1 and not args.no_extra and args.workload == "config3":
        try:
            extra = {"config2": extra_line(pkg, W, "config2", local, args, peak)}
        except Exception as e:                              # a secondary line must never cost the main one
            extra = {"config2": {"error": repr(e)}}
        try:
            extra["lbvh_build"] = lbvh_line(pkg, local, peak)
        except Exception as e:
            extra["lbvh_build"] = {"error": repr(e)}
    if rank == 0:
        out = {"metric": "volume-gather rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": cfg["label"], "photons": n_ph, "grid": cfg["grid"], "xres": cfg["xres"], "yres": cfg["yres"],
                          "stepsize": cfg["stepsize"], "nused": cfg["nused"], "maxdist": cfg["maxdist"], "photon_record_bytes": 144,
                          "l2_policy": "inputs_exceed_l2 (photon map %.1f GB)" % (n_ph * 160 / 1e9), "ray_order": "8x8 tiles",
                          "parallelism": "tiles/%d" % world},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),        # counted by the library at every launch site of the gather path (pv_launch_count): march x3, step keys, radix sort passes, cellgather, overflow pass, recurrence
               "clocks": clocks,
               "shoot": shoot, "frame": frame, "extra": extra,
               "build": {"seconds": build_s, "photons": n_ph, "photon_gen_host_s": gen_s, "note": "second pv_build of the resident set (warm, no allocation)"},
               "allgather": allgather,
               "lookups_per_s": stats.lookups * world / (total_ms * 1e-3) if world == 1 else None, "checksum_L": check}
        print(json.dumps(out), flush=True)
    pv.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
