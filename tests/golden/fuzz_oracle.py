#!/usr/bin/env python
"""Randomised cross-check of the CPU oracle against the REAL reference (this container only: needs oracle/_ref/ref_harness).

The committed goldens pin the oracle on fixed scenes; this script draws random ones -- medium kind and coefficients, phase
asymmetry, emission, one to three lights of random kind / position / cone, step size, volume integrator -- lets the unmodified
reference compute Li (and, for "photonvolume", shoot the photons first), and replays everything in the oracle with the
reference's MT19937 stream.  Bars as in tests/test_oracle_golden.py: T 1e-5, L 1e-5 relative, identical zero patterns, photon lists
identical in count and position.

    python tests/golden/fuzz_oracle.py [n_scenes=24] [first_seed=0]
"""
import json
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

load_package()
from cs348b_pbrt_b200 import sceneio, scenes  # noqa: E402
import oracle_lib as O  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def col(rng, lo, hi):
    return "[%g %g %g]" % tuple(rng.uniform(lo, hi, 3))


def random_scene(rng):
    kind = rng.choice(["homogeneous", "volumegrid", "exponential"])
    g = rng.uniform(-0.6, 0.8)
    common = '"color sigma_a" %s "color sigma_s" %s "color Le" %s "float g" [%g] "point p0" [-1 -1 -1] "point p1" [1 1 1]' % (
        col(rng, 0.05, 1.5), col(rng, 0.05, 2.0), col(rng, 0, 0.4) if rng.random() < 0.6 else "[0 0 0]", g)
    if kind == "homogeneous":
        vol = 'Volume "homogeneous" ' + common
    elif kind == "volumegrid":
        n = int(rng.choice([4, 9, 16]))
        dens = rng.uniform(0, 2, n * n * n) * (rng.random(n * n * n) < 0.8)
        vol = 'Volume "volumegrid" %s "integer nx" [%d] "integer ny" [%d] "integer nz" [%d] "float density" [%s]' % (
            common, n, n, n, " ".join("%.9g" % v for v in dens))
    else:
        up = rng.normal(size=3); up /= np.linalg.norm(up)
        vol = 'Volume "exponential" %s "float a" [%g] "float b" [%g] "vector updir" [%g %g %g]' % (common, rng.uniform(0.5, 3), rng.uniform(0.2, 2), *up)
    lights = []
    for _ in range(int(rng.integers(1, 4))):
        lk = rng.choice(["point", "spot", "distant"])
        if lk == "point":
            lights.append('LightSource "point" "point from" [%g %g %g] "color I" %s' % (*rng.uniform(-0.8, 0.8, 3), col(rng, 2, 30)))
        elif lk == "spot":
            lights.append('LightSource "spot" "point from" [%g %g %g] "point to" [%g %g %g] "color I" %s "float coneangle" [%g] "float conedeltaangle" [%g]' % (
                *rng.uniform(-0.8, 0.8, 3), *rng.uniform(-0.8, 0.8, 3), col(rng, 5, 60), rng.uniform(15, 70), rng.uniform(2, 14)))
        else:
            lights.append('LightSource "distant" "point from" [%g %g -3] "point to" [%g %g 0] "color L" %s' % (
                *rng.uniform(-1, 1, 2), *rng.uniform(-0.5, 0.5, 2), col(rng, 1, 8)))
    return vol, "\n".join(lights)


def random_area_quad(rng):
    """A DiffuseAreaLight over a randomly placed, randomly oriented parallelogram (two triangles) inside the box."""
    c = rng.uniform(-0.6, 0.6, 3); e1 = rng.normal(size=3) * 0.25; e2 = rng.normal(size=3) * 0.25
    P = [c - e1 - e2, c + e1 - e2, c + e1 + e2, c - e1 + e2]
    return ('AttributeBegin\nAreaLightSource "diffuse" "color L" %s\nShape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [%s]\nAttributeEnd'
            % (col(rng, 2, 12), "  ".join("%g %g %g" % tuple(q) for q in P)))


def random_glass(rng):
    """A glass object in the box: a (possibly partial, scaled, rotated) sphere or the wedge mesh of the all-maps scene; random index,
    dispersion (Vn = 0: none), reflectance / transmittance."""
    mat = 'Material "glass" "float index" [%g] "float Vn" [%g] "color Kr" %s "color Kt" %s' % (
        rng.uniform(1.1, 2.0), 0.0 if rng.random() < 0.4 else rng.uniform(1.5, 6.0), col(rng, 0, 1), col(rng, 0.2, 1))
    xf = "Translate %g %g %g\nRotate %g %g %g %g\nScale %g %g %g" % (*rng.uniform(-0.4, 0.4, 3), rng.uniform(0, 360), *rng.normal(size=3), *rng.uniform(0.6, 1.3, 3))
    if rng.random() < 0.5:
        extra = "" if rng.random() < 0.5 else ' "float zmin" [%g] "float zmax" [%g] "float phimax" [%g]' % (rng.uniform(-0.3, -0.05), rng.uniform(0.05, 0.3), rng.uniform(120, 360))
        shape = 'Shape "sphere" "float radius" [%g]%s' % (rng.uniform(0.2, 0.4), extra)
    else:
        shape = ('Scale 0.45 0.25 0.45\nShape "trianglemesh" "point P" [1 -1 -1  1 -1 1  -1 -1 1  -1 -1 -1  1 1 0  -1 1 0]\n'
                 '  "integer indices" [0 1 2  0 2 3  1 4 5  1 5 2  0 4 1  2 5 3  4 0 3  4 3 5]')
    return "AttributeBegin\n%s\n%s\n%s\nAttributeEnd" % (mat, xf, shape)


def second_volume(rng):
    """A second, overlapping Volume statement => AggregateVolume."""
    lo = rng.uniform(-1, 0.2, 3); hi = lo + rng.uniform(0.5, 1.2, 3)
    return ('Volume "homogeneous" "color sigma_a" %s "color sigma_s" %s "color Le" %s "float g" [%g] "point p0" [%g %g %g] "point p1" [%g %g %g]'
            % (col(rng, 0.05, 1.0), col(rng, 0.05, 1.5), col(rng, 0, 0.3), rng.uniform(-0.5, 0.7), *lo, *np.minimum(hi, 1.0)))


def relerr(a, b):
    return np.abs(np.asarray(a, np.float64) - b) / np.maximum(np.abs(np.asarray(b, np.float64)), 1e-30)


def scene_for_seed(seed, base=None):
    """The random scene of `seed`: .pbrt text, camera rays and parameters (everything main() and make_golden.py need)."""
    base = scenes.camera_rays(48, 48) if base is None else base
    rng = np.random.default_rng(1000 + seed)
    vol, lights = random_scene(rng)
    integ = ["single", "emission", "photonvolume"][seed % 3]
    variant = ["plain", "glass", "area", "aggregate"][(seed // 3) % 4]
    tail = ""
    if variant == "area":
        tail = random_area_quad(rng)
    elif variant == "aggregate":
        vol = vol + "\n" + second_volume(rng)
    elif variant == "glass":
        tail = random_glass(rng)
    stepsize = round(float(rng.uniform(0.03, 0.2)), 4)          # 4 decimals: the scene text (%g) and the oracle see the same number
    rays = base[np.sort(rng.choice(len(base), size=48, replace=False))].copy()
    rays["u_scatter"] = rng.random(len(rays)).astype(np.float32)
    rays["maxt"][:8] = rng.uniform(2.5, 4.5, 8).astype(np.float32)
    point = 'LightSource "point" "point from" [0 0.8 0] "color I" [20 20 20]'
    d = dict(vol=vol, lights=lights, integ=integ, variant=variant, stepsize=stepsize, rays=rays)
    if integ == "photonvolume":
        nused, maxdist, wanted, shoot_step = int(rng.integers(10, 60)), round(float(rng.uniform(0.15, 0.4)), 4), int(rng.integers(500, 1500)), round(float(rng.uniform(0.04, 0.12)), 4)
        text = scenes.cornell_pbrt(vol, wanted, stepsize=stepsize, nused=nused, maxdist=maxdist, shoot_step=shoot_step).replace(point, lights)
        d.update(nused=nused, maxdist=maxdist, wanted=wanted, shoot_step=shoot_step)
    else:
        text = scenes.volint_pbrt(integ, vol, stepsize=stepsize).replace(point, lights)
    d["text"] = text.replace("WorldEnd", tail + "\nWorldEnd")
    return d


def main():
    n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "all"], stdout=subprocess.DEVNULL)
    base = scenes.camera_rays(48, 48)
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(first, first + n_scenes):
            d = scene_for_seed(seed, base)
            vol, lights, integ, variant, stepsize, rays, text = d["vol"], d["lights"], d["integ"], d["variant"], d["stepsize"], d["rays"], d["text"]
            rf = os.path.join(tmp, "rays.bin"); sceneio.write_rays(rf, rays)
            scn = os.path.join(tmp, "s.scn"); out = os.path.join(tmp, "li.bin"); pho = os.path.join(tmp, "p.pho"); stats = os.path.join(tmp, "st.json")
            if integ == "photonvolume":
                nused, maxdist, wanted, shoot_step = d["nused"], d["maxdist"], d["wanted"], d["shoot_step"]
                ops = ["--shoot", "--dump-photons", pho, "--stats", stats, "--li", rf, "1000", out]
            else:
                ops = ["--vli", rf, "1000", out]
            side = os.path.join(tmp, "s.lights"); reg = os.path.join(tmp, "reg")
            export = {"plain": ["--export-scene", scn], "glass": ["--export-scene", scn], "area": ["--export-area-lights", scn, side], "aggregate": ["--export-regions", reg]}[variant]
            ops = export + ops
            f = os.path.join(tmp, "s.pbrt"); open(f, "w").write(text)
            r = subprocess.run([HARNESS, f] + ops, capture_output=True, text=True)
            if r.returncode != 0:
                print("seed %d: harness failed: %s" % (seed, r.stderr[-300:])); bad += 1; continue
            import contextlib
            ctx = contextlib.nullcontext()
            if variant == "aggregate":
                scene = sceneio.read_scene(reg + ".0.scn"); ctx = O.more_media(sceneio.read_scene(reg + ".1.scn"))
            else:
                scene = sceneio.read_scene(scn)
                if variant == "area":
                    ctx = O.area_lights(side)
            li = sceneio.read_spectra(out, b"PVLI0001", per=2); refL, refT = li[:, 0], li[:, 1]
            msg = []
            ctx.__enter__()
            if integ == "photonvolume":
                pos, wi, alpha = sceneio.read_photons(pho); st = json.load(open(stats))
                res = O.shoot(scene, wanted, shoot_step, stepsize, rng_mode=O.MT)
                if res["nshot"] != st["nshot"] or res["n"] != len(pos):
                    msg.append("photon list: nshot %d vs %d, n %d vs %d" % (res["nshot"], st["nshot"], res["n"], len(pos)))
                elif len(pos) and (np.abs(res["pos"] - pos).max() > 1e-5 or relerr(res["alpha"], alpha).max() > 1e-5):
                    msg.append("photon list: pos %.3g alpha %.3g" % (np.abs(res["pos"] - pos).max(), relerr(res["alpha"], alpha).max()))
                L, T = (refL * 0, refT * 0 + 1)
                if len(pos):
                    L, T, _ = O.gather(scene, O.KdTree(pos), wi, alpha, rays, stepsize, nused, maxdist, rng_mode=O.MT, mt_seed=1000)
                else:
                    L, T, _ = O.gather(scene, None, wi, alpha, rays, stepsize, nused, maxdist, rng_mode=O.MT, mt_seed=1000, flags=2)
            else:
                L, T, _ = O.volume_li(scene, rays, stepsize, O.SINGLE if integ == "single" else O.EMISSION, rng_mode=O.MT, mt_seed=1000)
            ctx.__exit__(None, None, None)
            ok = np.isfinite(refL) & np.isfinite(refT)
            if not np.array_equal(T[ok] == 0, refT[ok] == 0) or (refT[ok] > 0).any() and relerr(T[ok], refT[ok])[refT[ok] > 0].max() > 1e-5:
                zr = (T[ok] == 0) != (refT[ok] == 0)
                msg.append("T (zero-pattern differences: %d; max rel err elsewhere %.3g)" % (zr.sum(), relerr(T[ok], refT[ok])[(refT[ok] > 0) & (T[ok] > 0)].max() if ((refT[ok] > 0) & (T[ok] > 0)).any() else 0))
            m = ok & (refL > 0)
            if not np.array_equal(L[ok] == 0, refL[ok] == 0): msg.append("L zero pattern")
            if m.any() and relerr(L[m], refL[m]).max() > 1e-5: msg.append("L %.3g" % relerr(L[m], refL[m]).max())
            head = vol.split('"')[1]
            extra = " (%d photons, %d lit bins)" % (len(pos), int((refL > 0).sum())) if integ == "photonvolume" else " (%d lit bins)" % int((refL > 0).sum())
            print("seed %3d %-12s %-9s %-11s %d lights step %.4f: %s%s" % (seed, integ, variant, head, lights.count("LightSource"), stepsize,
                                                                    "ok" if not msg else "MISMATCH " + "; ".join(msg), extra))
            bad += bool(msg)
    print("%d of %d scenes disagree" % (bad, n_scenes))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
