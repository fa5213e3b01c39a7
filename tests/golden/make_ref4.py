"""BASELINE configs[3] at its SHIPPED SIZE: projectScene/pinkfloyd.pbrt exactly as the reference project ships it (5 M volume photons,
512 x 512, nused 500, maxdist .4, glass prism with dispersion, spot + point light, EXR output) with ONE change: 1 sample per pixel
instead of 32 (the unmodified reference needs 67 s of shooting + 136 s of rendering PER SAMPLE on this container's 8 cores; the photon
count, resolution and lookup size are what load the device path).  Copies the scene (and obj/prism.pbrt it includes) next to the
other test scenes, renders it with the unmodified reference (--ncores 1) into tests/golden/pinkfloyd_1spp_ref.npy (fp16, = the
decoded .exr) and measures the reference's own spread on other random streams (--ncores 2, 3, 5) into tests/golden/ref_spread.json.
Run in the container that has /root/reference:  python tests/golden/make_ref4.py   (about 15 minutes)"""
import json, os, shutil, subprocess, sys, tempfile, time
import numpy as np
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin_render import image_errors
REF = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref")
SRC = "/root/reference/projectScene"
name = "pinkfloyd_1spp"
text = open(os.path.join(SRC, "pinkfloyd.pbrt")).read()
assert '"integer pixelsamples" [32]' in text
text = text.replace('"integer pixelsamples" [32]', '"integer pixelsamples" [1]')
scenes_dir = os.path.join(ROOT, "tests", "scenes")
open(os.path.join(scenes_dir, name + ".pbrt"), "w").write(text)
os.makedirs(os.path.join(scenes_dir, "obj"), exist_ok=True)
shutil.copyfile(os.path.join(SRC, "obj", "prism.pbrt"), os.path.join(scenes_dir, "obj", "prism.pbrt"))
tmp = tempfile.mkdtemp()


def render(cores):
    t0 = time.time()
    subprocess.check_call([REF, "--ncores", str(cores), "--quiet", os.path.join(scenes_dir, name + ".pbrt")], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    bgra = cv2.imread(os.path.join(tmp, "pinkfloyd.exr"), cv2.IMREAD_UNCHANGED)
    print("ncores", cores, "%.0f s" % (time.time() - t0), flush=True)
    return bgra[..., [2, 1, 0]].astype(np.float16)


primary = render(1)
np.save(os.path.join(HERE, name + "_ref.npy"), primary)
runs = {}
for c in (2, 3, 5):
    e = image_errors(render(c).astype(np.float32), primary.astype(np.float32))
    runs["ncores_%d" % c] = {"e_mean": float(e[0]), "e_block": float(e[1])}
path = os.path.join(HERE, "ref_spread.json")
spread = json.load(open(path))
spread[name] = {"runs": runs, "e_mean": max(r["e_mean"] for r in runs.values()), "e_block": max(r["e_block"] for r in runs.values())}
json.dump(spread, open(path, "w"), indent=1, sort_keys=True)
print(name, spread[name]["e_mean"], spread[name]["e_block"])
