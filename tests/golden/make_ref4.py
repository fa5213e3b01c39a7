"""BASELINE configs[3] at its SHIPPED SIZE: projectScene/pinkfloyd.pbrt exactly as the reference project ships it (5 M volume photons,
512 x 512, nused 500, maxdist .4, glass prism with dispersion, spot + point light, EXR output) with ONE change: 1 sample per pixel
instead of 32 (the unmodified reference needs 67 s of shooting + 136 s of rendering PER SAMPLE on this container's 8 cores; the photon
count, resolution and lookup size are what load the device path).  Renders it with the unmodified reference (--ncores 1) into tests/golden/pinkfloyd_1spp_ref.npy (fp16, = the
decoded .exr) and measures the reference's own spread on other random streams into tests/golden/ref_spread.json.  For a 512 x 512
frame the reference cuts the image into max(32 * ncores, pixels / 256) = 1024 tasks for any ncores <= 32, and a task's RNG seed and
sample window depend on the task number alone -- so --ncores 2, 3, 5 re-draw the PHOTONS but render with the very same random
numbers (light choice per march step, roulette, tau offsets); the single-scattered light of the 0.8 degree spot beam, which a 0.05
march step hits or misses, is the noisiest part of this image and only shows in runs with another task count: --ncores 40, 70, 130
and 260 (2048 ... 16384 tasks) are therefore part of the spread.  Run in the container that has /root/reference:
python tests/golden/make_ref4.py   (about an hour on 8 cores; SPREAD_ONLY=130,260 adds runs to an existing golden)"""
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
tmp = tempfile.mkdtemp()
scenes_dir = tmp                                  # the 1-spp variant lives in the scratch directory only (the test makes its own)
open(os.path.join(scenes_dir, name + ".pbrt"), "w").write(text)
shutil.copytree(os.path.join(SRC, "obj"), os.path.join(scenes_dir, "obj"))


def render(cores):
    t0 = time.time()
    subprocess.check_call([REF, "--ncores", str(cores), "--quiet", os.path.join(scenes_dir, name + ".pbrt")], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    bgra = cv2.imread(os.path.join(tmp, "pinkfloyd.exr"), cv2.IMREAD_UNCHANGED)
    print("ncores", cores, "%.0f s" % (time.time() - t0), flush=True)
    return bgra[..., [2, 1, 0]].astype(np.float16)


path = os.path.join(HERE, "ref_spread.json")
spread = json.load(open(path))
if os.environ.get("SPREAD_ONLY"):
    primary = np.load(os.path.join(HERE, name + "_ref.npy"))
    runs = spread[name]["runs"]
    cores = [int(c) for c in os.environ["SPREAD_ONLY"].split(",")]
else:
    primary = render(1)
    np.save(os.path.join(HERE, name + "_ref.npy"), primary)
    runs = {}
    cores = [2, 3, 5, 40, 70, 130, 260]
for c in cores:
    e = image_errors(render(c).astype(np.float32), primary.astype(np.float32))
    runs["ncores_%d" % c] = {"e_mean": float(e[0]), "e_block": float(e[1])}
spread[name] = {"runs": runs, "e_mean": max(r["e_mean"] for r in runs.values()), "e_block": max(r["e_block"] for r in runs.values())}
json.dump(spread, open(path, "w"), indent=1, sort_keys=True)
print(name, spread[name]["e_mean"], spread[name]["e_block"])
