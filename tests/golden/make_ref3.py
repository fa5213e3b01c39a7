"""Round-2 addition to make_golden.py / make_ref2.py: the scene that takes photonmap.cpp:308 (LPhoton on the INDIRECT map, final
gathering off) next to :179 (LPhoton on the caustic map) -- cornell_surf_nofg_e2e: every photon map, glass wedge, 72 x 72, 4 spp.
Writes tests/scenes/cornell_surf_nofg_e2e.pbrt, the unmodified reference's render of it (--ncores 1) as
tests/golden/cornell_surf_nofg_e2e_ref.npy and its run-to-run spread (--ncores 2, 3, 5) into tests/golden/ref_spread.json, leaving
the other scenes' entries as they are.  Run in the container that has /root/reference:  python tests/golden/make_ref3.py"""
import json, os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin_render import read_pfm, image_errors
from __graft_entry__ import load_package
load_package()
from cs348b_pbrt_b200 import scenes
REF = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref")
name = "cornell_surf_nofg_e2e"
text = scenes.cornell_surf_pbrt(nphotons=20000, caustic=5000, indirect=20000, finalgather=False, xres=72, yres=72,
                                outfile=name + ".pfm").replace('"integer pixelsamples" [1]', '"integer pixelsamples" [4]')
scene = os.path.join(ROOT, "tests", "scenes", name + ".pbrt")
open(scene, "w").write(text)
tmp = tempfile.mkdtemp()
run = lambda c: subprocess.check_call([REF, "--ncores", str(c), "--quiet", scene], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
run(1)
primary = read_pfm(os.path.join(tmp, name + ".pfm")).astype(np.float16)
np.save(os.path.join(HERE, name + "_ref.npy"), primary)
runs = {}
for c in (2, 3, 5):
    run(c)
    e = image_errors(read_pfm(os.path.join(tmp, name + ".pfm")).astype(np.float16).astype(np.float32), primary.astype(np.float32))
    runs["ncores_%d" % c] = {"e_mean": float(e[0]), "e_block": float(e[1])}
path = os.path.join(HERE, "ref_spread.json")
spread = json.load(open(path))
spread[name] = {"runs": runs, "e_mean": max(r["e_mean"] for r in runs.values()), "e_block": max(r["e_block"] for r in runs.values())}
json.dump(spread, open(path, "w"), indent=1, sort_keys=True)
print(name, spread[name]["e_mean"], spread[name]["e_block"])
