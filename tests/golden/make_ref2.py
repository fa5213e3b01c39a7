"""The reference's own run-to-run spread, per end-to-end scene: the unmodified reference renders every scene of tests/scenes/ again
with OTHER task counts (--ncores 2, 3, 5 instead of 1: it seeds its RNGs and scrambles its samples per task, so these are the same
renderer on other random streams) and the distances of those renders from the primary one (tests/golden/<name>_ref.npy) are
written to tests/golden/ref_spread.json.  The drop-in's whole-image
tolerance is stated against these numbers (tests/test_dropin_render.py).  Also renders BASELINE config 1 VERBATIM
(projectScene/volumescene_png.pbrt as shipped: PNG, 300 x 300).  Run in the container that has /root/reference:
    python tests/golden/make_ref2.py
"""
import json, os, shutil, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin_render import read_pfm, read_png_rgb8, image_errors
REF = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref")
CORES = (2, 3, 5)
tmp = tempfile.mkdtemp()
spread = {}
# an area light on the path (DiffuseAreaLight over a quad next to the point light, photonvolume integrator): scene + primary render
from __graft_entry__ import load_package
load_package()
from cs348b_pbrt_b200 import scenes
text = scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 20000, xres=64, yres=64, outfile="cornell_area_e2e.pfm").replace("WorldEnd", scenes.AREA_QUAD + "\nWorldEnd")
open(os.path.join(ROOT, "tests", "scenes", "cornell_area_e2e.pbrt"), "w").write(text)
subprocess.check_call([REF, "--ncores", "1", "--quiet", os.path.join(ROOT, "tests", "scenes", "cornell_area_e2e.pbrt")], cwd=tmp,
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
np.save(os.path.join(HERE, "cornell_area_e2e_ref.npy"), read_pfm(os.path.join(tmp, "cornell_area_e2e.pfm")).astype(np.float16))
for name in ("config1_volumescene", "config4_prism", "cornell_surf_e2e", "sphere_e2e", "cornell_e2e", "volint_single_e2e", "volint_emission_e2e",
             "cornell_area_e2e"):
    primary = np.load(os.path.join(HERE, name + "_ref.npy")).astype(np.float32)
    runs = {}
    for c in CORES:
        subprocess.check_call([REF, "--ncores", str(c), "--quiet", os.path.join(ROOT, "tests", "scenes", name + ".pbrt")], cwd=tmp,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        img = read_pfm(os.path.join(tmp, name + ".pfm")).astype(np.float16)
        e = image_errors(img.astype(np.float32), primary)
        runs["ncores_%d" % c] = {"e_mean": float(e[0]), "e_block": float(e[1])}
    spread[name] = {"runs": runs, "e_mean": max(r["e_mean"] for r in runs.values()), "e_block": max(r["e_block"] for r in runs.values())}
    print(name, spread[name]["e_mean"], spread[name]["e_block"], flush=True)
# config 1 verbatim: the scene file is the reference project's own data file, rendered where it lies (the GPU test reads the copy
# __graft_entry__.build() stages under the git-ignored baseline/_ref/projectScene)
dst = "/root/reference/projectScene/volumescene_png.pbrt"
lin = lambda path: (read_png_rgb8(path).astype(np.float32) / 255.0) ** 2.2
subprocess.check_call([REF, "--ncores", "1", "--quiet", dst], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
shutil.copyfile(os.path.join(tmp, "volume.png"), os.path.join(HERE, "volumescene_png_ref.png"))
primary = lin(os.path.join(HERE, "volumescene_png_ref.png"))
runs = {}
for c in CORES:
    subprocess.check_call([REF, "--ncores", str(c), "--quiet", dst], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    e = image_errors(lin(os.path.join(tmp, "volume.png")), primary)
    runs["ncores_%d" % c] = {"e_mean": float(e[0]), "e_block": float(e[1])}
spread["volumescene_png"] = {"runs": runs, "e_mean": max(r["e_mean"] for r in runs.values()), "e_block": max(r["e_block"] for r in runs.values())}
print("volumescene_png", spread["volumescene_png"]["e_mean"], spread["volumescene_png"]["e_block"])
json.dump(spread, open(os.path.join(HERE, "ref_spread.json"), "w"), indent=1, sort_keys=True)
