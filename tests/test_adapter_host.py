"""CPU: the HOST logic of the drop-in binary (cs348b-pbrt_b200/host/pv_pbrt_adapter.cpp linked into the reference renderer,
baseline/_ref/pbrt_b200) run against a TEST DOUBLE of libpv.so (tests/mock/mock_libpv.c, put in front of the real library with
LD_LIBRARY_PATH).  The double returns a hash of each ray's global stream index as "radiance" and logs every call, so these
tests see what the adapter asks the device to do -- call order, which device gets which rays under which index, how results
are stitched -- without a GPU.  Nothing here says anything about the CUDA path's numbers; that is tests/test_gpu_*.py."""
import os
import re
import shutil
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "baseline", "_ref", "pbrt_b200")
pytestmark = pytest.mark.skipif(not os.path.exists(BIN) or shutil.which("gcc") is None,
                                reason="baseline/_ref/pbrt_b200 not built (needs the reference sources: make -C cs348b-pbrt_b200/host)")


@pytest.fixture(scope="module")
def mock(tmp_path_factory):
    d = tmp_path_factory.mktemp("mockpv")
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "mock", "mock_libpv.c"), "-o", str(d / "libpv.so")])
    return d


def render(mock, tmp_path, name, text, devices=None):
    scene = tmp_path / (name + ".pbrt")
    scene.write_text(text)
    log = tmp_path / (name + ".log")
    env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log))
    env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
    if devices:
        env["PV_DEVICES"] = devices
    out = subprocess.run([BIN, "--quiet", "--ncores", "2", str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    img = (tmp_path / (name + ".pfm")).read_bytes()
    return img, log.read_text().splitlines(), out.stderr


def calls(log, what):
    return [dict((k, int(v)) for k, v in re.findall(r"(\w+)=(\d+)", l)) for l in log if l.startswith(what + " ")]


def test_frame_is_sharded_over_devices_by_global_ray_index(mock, tmp_path, pkg):
    """VolumeIntegrator "single", PV_DEVICES=0,1,2: the scene goes to three contexts, the camera rays of the frame to three
    concurrent pv_volume_li calls that partition [0, n) in order, each with its first ray's global index as stream base; the
    stitched image is byte-identical to the one-device image."""
    from cs348b_pbrt_b200 import scenes
    vol = scenes.VOLINT_MEDIA["volint_homog"][0]
    one, log1, _ = render(mock, tmp_path, "one", scenes.volint_pbrt("single", vol, xres=160, yres=160, outfile="one.pfm"))
    three, log3, err = render(mock, tmp_path, "three", scenes.volint_pbrt("single", vol, xres=160, yres=160, outfile="three.pfm"), "0,1,2")
    assert one[:20] == three[:20] and one == three
    assert [c["dev"] for c in calls(log1, "create")] == [0] and len(calls(log1, "volume_li_single")) == 1
    assert [c["dev"] for c in calls(log3, "create")] == [0, 1, 2]
    assert sorted(c["dev"] for c in calls(log3, "set_scene")) == [0, 1, 2]
    li = sorted(calls(log3, "volume_li_single"), key=lambda c: c["base"])
    n = calls(log1, "volume_li_single")[0]["n"]
    assert [c["dev"] for c in li] == [0, 1, 2]
    assert li[0]["base"] == 0 and all(a["base"] + a["n"] == b["base"] for a, b in zip(li, li[1:])) and li[-1]["base"] + li[-1]["n"] == n
    assert "replicated on 2 more device(s)" in err


def test_photon_map_is_replicated_before_the_frame_is_sharded(mock, tmp_path, pkg):
    """VolumeIntegrator "photonvolume", PV_DEVICES=2,0: photons are shot and the map built on the FIRST listed device, then
    broadcast device to device (one communicator over both contexts, pv_broadcast_photons -- no read-back through host
    memory) and built on the other, then the frame is split between the two."""
    from cs348b_pbrt_b200 import scenes
    text = lambda out: scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 1000, xres=128, yres=128, outfile=out)
    one, log1, _ = render(mock, tmp_path, "pone", text("pone.pfm"))
    two, log2, _ = render(mock, tmp_path, "ptwo", text("ptwo.pfm"), "2,0")
    assert one == two
    assert [c["dev"] for c in calls(log2, "create")] == [2, 0]
    assert [c["dev"] for c in calls(log2, "shoot")] == [2]
    order = [l.split()[0] + ":" + re.search(r"dev[s]?=(\d+)", l).group(1) for l in log2]
    assert order.index("build:2") < order.index("comm_init_all:2") < order.index("broadcast_photons:2") < order.index("build:0") < order.index("gather:2")
    assert not calls(log2, "get_photons") and not calls(log2, "set_photons")            # nothing staged through the host
    b = calls(log2, "broadcast_photons")
    assert len(b) == 1 and b[0]["n"] == 1000 and b[0]["to"] == 1
    assert [c["n"] for c in calls(log2, "build")] == [1000, 1000]
    g = sorted(calls(log2, "gather"), key=lambda c: c["base"])
    assert [c["dev"] for c in g] == [2, 0] and g[0]["base"] == 0 and g[0]["n"] == g[1]["base"]
    assert g[1]["base"] + g[1]["n"] == calls(log1, "gather")[0]["n"]


def test_emission_scene_off_the_path_goes_down_with_its_medium_only(mock, tmp_path, pkg):
    """pbrt's default volume integrator on a scene whose surfaces / lights the device path does not know (a disk-shaped area
    light): the adapter exports the medium alone (no primitives, no lights) and the frame is one pv_volume_li(emission)."""
    from cs348b_pbrt_b200 import scenes
    _, log, _ = render(mock, tmp_path, "volint", scenes.volint_offpath_pbrt())
    s = calls(log, "set_scene")
    assert len(s) == 1 and s[0]["prims"] == 0 and s[0]["lights"] == 0
    assert len(calls(log, "volume_li_emission")) == 1 and not calls(log, "volume_li_single") and not calls(log, "gather")


def test_secondary_rays_are_batched_across_render_threads_without_loss(mock, tmp_path, pkg):
    """The glass wedge's specular bounces reach VolumeIntegrator::Li one ray at a time on every render thread; the adapter
    gathers them into batches (one device call each).  Every such ray is in exactly one batch: the rays the adapter counted
    equal the rays the device calls carried, batches hold at most 4096 rays, and the frame itself is still one call."""
    from cs348b_pbrt_b200 import scenes
    text = scenes.volint_e2e_pbrt("single", scenes.VOLINT_MEDIA["volint_homog"][0], xres=96, yres=96, spp=2, outfile="wedge.pfm")
    _, log, err = render(mock, tmp_path, "wedge", text)
    m = re.search(r"volume term of (\d+) secondary rays \(specular bounces\) in (\d+) batched device calls", err)
    assert m, err[-1000:]
    nsec, nbatches = int(m.group(1)), int(m.group(2))
    li = calls(log, "volume_li_single")
    frame = [c for c in li if c["base"] == 0]
    assert len(frame) == 1 and frame[0]["n"] >= 96 * 96 * 2          # every camera sample of the frame (the film's sample extent)
    small = [c for c in li if c is not frame[0]]
    assert nsec > 1000 and len(small) == nbatches and sum(c["n"] for c in small) == nsec
    assert max(c["n"] for c in small) <= 4096
    assert all(c["base"] >> 63 for c in small)               # secondary rays live in the upper half of the stream-index space


def test_final_gather_wavefront_is_reproducible_across_thread_counts(mock, tmp_path, pkg):
    """All photon maps + final gathering (no specular surfaces): Preprocess is ONE pv_shoot_maps followed by the reads the untouched
    PhotonIntegrator needs (caustic / indirect lists, radiance photons), the final-gather rays of primary hits go down as
    pv_final_gather batches on a second context whose grid holds the radiance photons, in task order with consecutive global
    indices -- so the frame does not depend on how many render threads produced it (1 vs 2 cores, same task count)."""
    from cs348b_pbrt_b200 import scenes
    def text(out):
        t = scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 1000, xres=128, yres=128, outfile=out)
        t = t.replace('"bool finalgather" ["false"]', '"bool finalgather" ["true"] "integer finalgathersamples" [4]')
        return t.replace('"integer indirectphotons" [0]', '"integer indirectphotons" [500]')
    imgs, logs, errs = [], [], []
    for cores in ("1", "2"):
        scene = tmp_path / ("fg%s.pbrt" % cores); scene.write_text(text("fg%s.pfm" % cores))
        log = tmp_path / ("fg%s.log" % cores)
        env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log))
        env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
        out = subprocess.run([BIN, "--quiet", "--ncores", cores, str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        imgs.append((tmp_path / ("fg%s.pfm" % cores)).read_bytes()); logs.append(log.read_text().splitlines()); errs.append(out.stderr)
    assert imgs[0] == imgs[1]
    log, err = logs[1], errs[1]
    names = [l.split()[0] for l in log]
    # Preprocess: one pass for every map, then the lists the CPU surface integrator reads, radiance photons, the volume grid
    assert names[:3] == ["create", "set_scene", "shoot_maps"] and names.count("shoot_maps") == 1 and "shoot" not in names
    assert [c["map"] for c in calls(log, "get_map_photons")] == [2, 4]            # indirect (caustic map empty: not read), radiance sites
    assert names.index("radiance_photons") < names.index("build")
    # the final-gather context: scene + radiance photons + their grid, before the first batch
    assert names.index("select_map") < names.index("final_gather") and calls(log, "select_map")[0]["map"] == 4
    fg = calls(log, "final_gather")
    m = re.search(r"final gathering of primary hits on the GPU: (\d+) gather rays", err)
    assert m and sum(c["n"] for c in fg) == int(m.group(1)) > 10000
    assert fg[0]["base"] == 0 and all(a["base"] + a["n"] == b["base"] for a, b in zip(fg, fg[1:]))
    assert len(calls(log, "gather")) == 1                                        # the volume term of the frame: still one call


def test_material_off_the_path_is_refused_not_approximated(mock, tmp_path, pkg):
    """A scene whose materials the device description cannot hold (here "plastic") must not get photon maps traced with a
    stand-in material: the drop-in stops with an error naming the reason, before any photon is shot."""
    from cs348b_pbrt_b200 import scenes
    text = scenes.cornell_surf_pbrt(nphotons=500, caustic=200, indirect=500, finalgather=False, xres=32, yres=32, outfile="plastic.pfm")
    assert 'Material "matte" "color Kd" [.6 .6 .6]' in text
    text = text.replace('Material "matte" "color Kd" [.6 .6 .6]', 'Material "plastic" "color Kd" [.6 .6 .6] "color Ks" [.3 .3 .3]', 1)
    scene = tmp_path / "plastic.pbrt"; scene.write_text(text)
    log = tmp_path / "plastic.log"
    env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log))
    out = subprocess.run([BIN, "--quiet", "--ncores", "2", str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode != 0
    assert "neither matte nor glass" in out.stderr
    lines = log.read_text().splitlines() if log.exists() else []
    assert not [l for l in lines if l.startswith("shoot")]


@pytest.mark.parametrize("accel", ["kdtree", "grid", "bvh+PV_BVH=gpu"])
def test_scene_without_a_reference_bvh_gets_one_from_the_device(mock, tmp_path, pkg, accel):
    """Accelerator "kdtree" (pbrt's own default) and "grid" hold no LinearBVHNode array to export: the adapter hands the bounds of
    the refined primitives to pv_build_bvh (max 4 per leaf, the reference's "maxnodeprims" default), and describes the scene to the
    device with the primitives IN THE ORDER THE BUILDER RETURNED and the builder's node array.  PV_BVH=gpu asks for the same with
    the "bvh" accelerator.  The double returns a one-leaf tree, with the primitives reversed under MOCK_PV_BVH_REVERSE=1."""
    from cs348b_pbrt_b200 import scenes
    vol = scenes.VOLINT_MEDIA["volint_homog"][0]
    text = scenes.volint_pbrt("single", vol, xres=32, yres=32, outfile="acc.pfm")
    assert "WorldBegin" in text and "Accelerator" not in text
    env_extra = {}
    if accel.startswith("bvh"):
        env_extra["PV_BVH"] = "gpu"
    else:
        text = text.replace("WorldBegin", 'Accelerator "%s"\nWorldBegin' % accel, 1)
    orders = []
    for rev in ("0", "1"):
        scene = tmp_path / ("acc%s.pbrt" % rev); scene.write_text(text)
        log = tmp_path / ("acc%s.log" % rev)
        env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log), MOCK_PV_BVH_REVERSE=rev, **env_extra)
        env.pop("PV_DEVICES", None)
        out = subprocess.run([BIN, "--quiet", "--ncores", "2", str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = log.read_text().splitlines()
        b = calls(lines, "build_bvh")
        assert len(b) == 1 and b[0]["n"] == 12 and b[0]["maxprims"] == 4 and b[0]["cap"] >= 2 * 12 - 1
        assert "scene BVH built on the GPU: 12 primitives -> 1 nodes" in out.stderr
        s = [l for l in lines if l.startswith("set_scene ")]
        assert len(s) == 1 and calls(lines, "set_scene")[0]["prims"] == 12 and calls(lines, "set_scene")[0]["nodes"] == 1
        assert lines.index(s[0]) > lines.index([l for l in lines if l.startswith("build_bvh ")][0])
        orders.append(s[0].split("order:")[1].split(","))
    assert len(orders[0]) == 12 and orders[0] == orders[1][::-1] and orders[0] != orders[1]


def test_the_reference_bvh_is_exported_when_there_is_one(mock, tmp_path, pkg):
    """default route: Accelerator "bvh" -> the reference's own node array, no device build"""
    from cs348b_pbrt_b200 import scenes
    vol = scenes.VOLINT_MEDIA["volint_homog"][0]
    _, log, err = render(mock, tmp_path, "refbvh", scenes.volint_pbrt("single", vol, xres=32, yres=32, outfile="refbvh.pfm"))
    assert not calls(log, "build_bvh") and calls(log, "set_scene")[0]["nodes"] > 1


@pytest.mark.parametrize("final_gather", [True, False])
def test_lphoton_of_primary_hits_goes_to_the_device(mock, tmp_path, pkg, final_gather):
    """photonmap.cpp:179 (caustic term) and, with final gathering off, :308 (indirect term): the LPhoton lookups of the primary hits
    of a group of tasks go down as ONE pv_surface_lphoton per map, on a context of their own that holds that map and its grid
    (k = the integrator's nused, path count = the shooter's for that map), and the sums come back multiplied with the surface's
    rho / pi.  The clone of the integrator that shades primary hits no longer sees those maps (no CPU kd-tree lookup is left for
    them); the frame is the same for 1 and 2 render threads."""
    from cs348b_pbrt_b200 import scenes
    def text(out):
        t = scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 1000, xres=96, yres=96, outfile=out)
        if final_gather:
            t = t.replace('"bool finalgather" ["false"]', '"bool finalgather" ["true"] "integer finalgathersamples" [4]')
        t = t.replace('"integer indirectphotons" [0]', '"integer indirectphotons" [500]')
        assert '"integer causticphotons" [0]' in t
        return t.replace('"integer causticphotons" [0]', '"integer causticphotons" [300]')
    imgs, logs, errs = [], [], []
    for cores in ("1", "2"):
        name = "lp%s%d" % (cores, final_gather)
        scene = tmp_path / (name + ".pbrt"); scene.write_text(text(name + ".pfm"))
        log = tmp_path / (name + ".log")
        env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log))
        env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
        out = subprocess.run([BIN, "--quiet", "--ncores", cores, str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        imgs.append((tmp_path / (name + ".pfm")).read_bytes()); logs.append(log.read_text().splitlines()); errs.append(out.stderr)
    assert imgs[0] == imgs[1]
    log, err = logs[1], errs[1]
    maps = [1] if final_gather else [1, 2]
    sm = [c for c in calls(log, "set_map_photons") if c["map"] in (1, 2)]
    assert [c["map"] for c in sm] == maps and [c["n"] for c in sm] == ([300] if final_gather else [300, 500])
    lp = calls(log, "surface_lphoton")
    assert sorted(set(c["map"] for c in lp)) == maps
    assert all(c["k"] == 50 and c["paths"] == 4096 for c in lp)                  # nused default 50; the double's path count
    per_map = [sum(c["n"] for c in lp if c["map"] == m) for m in maps]
    m = re.search(r"LPhoton of primary hits on the GPU \((.*) map\): (\d+) lookups", err)
    assert m and m.group(1) == ("caustic" if final_gather else "caustic + indirect") and int(m.group(2)) == sum(per_map)
    assert per_map[0] > 2000 and len(set(per_map)) == 1                          # one lookup per primary hit on a diffuse surface, per map
    assert bool(calls(log, "final_gather")) == final_gather


def test_direct_lighting_of_primary_hits_goes_to_the_device(mock, tmp_path, pkg):
    """UniformSampleAllLights / EstimateDirect for delta lights (core/integrator.cpp:47-79, 137-163) at primary hits: the light sample,
    BSDF value and cosine / pdf factor are the reference's own code per hit; the shadow rays of a group of render tasks and their
    transmittance through the medium (VolumeIntegrator::Transmittance with sample == NULL: step = 4 * stepsize, a random offset
    per ray) are ONE pv_occluded + ONE pv_transmittance.  Same frame for 1 and 2 render threads.  With an area light in the scene
    the reference's own per-ray direct lighting stays (its BSDF-sampling half needs the identity of the surface a ray hits)."""
    from cs348b_pbrt_b200 import scenes
    def text(out, area=False):
        t = scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 1000, xres=96, yres=96, outfile=out)
        t = t.replace('"integer causticphotons" [0]', '"integer causticphotons" [300]')
        return t.replace("WorldEnd", scenes.AREA_QUAD + "\nWorldEnd") if area else t
    imgs, logs, errs = [], [], []
    for name, cores, area in (("dl1", "1", False), ("dl2", "2", False), ("dla", "2", True)):
        scene = tmp_path / (name + ".pbrt"); scene.write_text(text(name + ".pfm", area))
        log = tmp_path / (name + ".log")
        env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log))
        env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
        out = subprocess.run([BIN, "--quiet", "--ncores", cores, str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        imgs.append((tmp_path / (name + ".pfm")).read_bytes()); logs.append(log.read_text().splitlines()); errs.append(out.stderr)
    assert imgs[0] == imgs[1]
    log, err = logs[1], errs[1]
    occ, tr = calls(log, "occluded"), calls(log, "transmittance")
    assert len(occ) == len(tr) >= 1 and [c["n"] for c in occ] == [c["n"] for c in tr]
    assert all(c["step1000"] == 200 and c["u"] == 1 for c in tr)                  # 4 * stepsize (0.05); offsets in [0, 1)
    m = re.search(r"direct lighting of primary hits on the GPU: (\d+) shadow rays", err)
    assert m and int(m.group(1)) == sum(c["n"] for c in occ) > 2000               # one per lit diffuse primary hit and light
    names = [l.split()[0] for l in log]
    assert names.index("occluded") > names.index("build") and names.index("occluded") < names.index("gather")
    # area light in the scene: no device direct lighting, the other device terms stay
    assert not calls(logs[2], "occluded") and "direct lighting of primary hits" not in errs[2]
    assert calls(logs[2], "surface_lphoton")


def test_surface_terms_behind_specular_bounces_are_queued_too(mock, tmp_path, pkg):
    """Glass wedge in the box, every photon map, final gathering: a camera ray that meets the glass is followed by the reference's
    SpecularReflect / SpecularTransmit, and the hit BEHIND the bounce is shaded like a primary hit -- its shadow rays, LPhoton
    lookups and final-gather rays join the same device batches, their weights carrying the throughput of the bounce (BSDF factor x
    volume transmittance of the secondary ray).  The secondary rays themselves are deferred as well: their volume terms are ONE
    pv_gather_indexed call per group, and the transmittances it returns are folded into the queued weights afterwards, chain by chain
    (checked against the previous, recursive evaluation of the same frame: equal to 1e-6).  So the batches hold MORE work
    than with "maxspeculardepth 1" (no bounce is followed), and the frame is the same for 1 and 2 render threads."""
    from cs348b_pbrt_b200 import scenes
    def text(out, depth=None):
        t = scenes.cornell_surf_pbrt(nphotons=1000, caustic=300, indirect=500, finalgather=True, fgsamples=4, xres=96, yres=96, outfile=out)   # 96 x 96: the same 64 tasks for 1 and 2 cores
        if depth is not None:
            assert 'SurfaceIntegrator "photonmap"' in t
            t = t.replace('SurfaceIntegrator "photonmap"', 'SurfaceIntegrator "photonmap" "integer maxspeculardepth" [%d]' % depth, 1)
        return t
    res = {}
    for name, cores, depth in (("sp1", "1", None), ("sp2", "2", None), ("sp0", "2", 1)):
        scene = tmp_path / (name + ".pbrt"); scene.write_text(text(name + ".pfm", depth))
        log = tmp_path / (name + ".log")
        env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log))
        env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
        out = subprocess.run([BIN, "--quiet", "--ncores", cores, str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = log.read_text().splitlines()
        res[name] = dict(img=(tmp_path / (name + ".pfm")).read_bytes(), err=out.stderr,
                         shadows=sum(c["n"] for c in calls(lines, "occluded")), lookups=sum(c["n"] for c in calls(lines, "surface_lphoton")),
                         fg=sum(c["n"] for c in calls(lines, "final_gather")), secondary=len([c for c in calls(lines, "gather") if c.get("indexed")]))
    assert res["sp1"]["img"] == res["sp2"]["img"]
    # volume term of the rays behind the bounces: ONE indexed device call per group of render tasks (here one group), not one per ray
    assert res["sp2"]["secondary"] == 1 and res["sp0"]["secondary"] == 0
    m = re.search(r"volume term of (\d+) rays behind specular bounces: 1 device call", res["sp2"]["err"])
    assert m and int(m.group(1)) > 1000
    for k in ("shadows", "lookups", "fg"):
        assert res["sp2"][k] > res["sp0"][k] > 0, k                               # the hits behind the glass queued their terms as well
        assert res["sp1"][k] == res["sp2"][k]


PROJECT_SCENES = "/root/reference/projectScene"


@pytest.mark.skipif(not os.path.isdir(PROJECT_SCENES), reason="the reference's data files are not here (this container only)")
def test_every_scene_the_project_ships_is_accepted_and_takes_the_device_routes(mock, tmp_path, pkg):
    """All ten .pbrt files of the reference's projectScene/ (read from the reference tree, resolution cut to 64 x 64 and 1 spp so the
    CPU side takes a second; photon counts, integrators, lights, materials, shapes, includes and output formats as shipped) run
    through the drop-in against the double: none is refused, every one shoots its photons with ONE device call (all maps when
    the scene asks for caustic photons, volume-only else), builds the volume map, sends the frame's volume term down as one gather
    and its shadow rays / caustic lookups as device batches; the files that name .exr / .png outputs write them."""
    import glob, shutil
    work = tmp_path / "proj"; work.mkdir()
    for sub in ("obj", "textures"):
        shutil.copytree(os.path.join(PROJECT_SCENES, sub), work / sub)
    files = sorted(glob.glob(os.path.join(PROJECT_SCENES, "*.pbrt")))
    assert len(files) == 10
    for f in files:
        name = os.path.basename(f)[:-5]
        text = open(f).read()
        text = re.sub(r'"integer pixelsamples" \[\d+\]', '"integer pixelsamples" [1]', text)
        text = re.sub(r'"integer ([xy])resolution" \[\d+\]', r'"integer \1resolution" [64]', text)
        (work / (name + ".pbrt")).write_text(text)
        log = work / (name + ".log")
        env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log))
        env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
        out = subprocess.run([BIN, "--quiet", "--ncores", "4", name + ".pbrt"], cwd=work, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, (name, out.stderr[-1500:])
        lines = log.read_text().splitlines()
        names = [l.split()[0] for l in lines]
        wants_surface_maps = int(re.search(r'"integer causticphotons"\s*\[?\s*(\d+)', text).group(1)) > 0 if "causticphotons" in text else True
        assert names.count("shoot_maps" if wants_surface_maps else "shoot") == 1 and "Shooting photons" not in out.stderr, name
        assert names.count("build") == 1 and names.count("set_scene") == 1, name
        frame = [c for c in calls(lines, "gather") if not c.get("indexed")]
        assert len(frame) == 1 and frame[0]["n"] >= 64 * 64, name                 # the camera rays of the frame: one call
        outfile = re.search(r'"string filename"\s+"([^"]+)"', text).group(1)
        assert (work / outfile).exists() and (work / outfile).stat().st_size > 1000, (name, outfile)
        if outfile.endswith(".exr"):
            assert (work / outfile).read_bytes()[:4] == b"\x76\x2f\x31\x01"       # OpenEXR magic


def test_film_accumulation_is_the_same_for_any_thread_count(mock, tmp_path, pkg):
    """The frame's samples are added to the film by the task pool, blocks of tiles at least as wide as the pixel filter reaches, one
    colour of a 2 x 2 checkerboard of blocks per round (no two threads of a round can touch the same pixel, every pixel gets its
    contributions in an order fixed by the tile layout).  A gaussian filter 9 pixels wide over 16 x 16-pixel tiles needs blocks of
    2 x 2 tiles: byte-identical frames from 1 and 2 render threads (128 x 128 image: 64 tasks in both runs)."""
    from cs348b_pbrt_b200 import scenes
    imgs = []
    for cores in ("1", "2"):
        name = "film" + cores
        t = scenes.cornell_pbrt(scenes.HOMOG_VOLUME, 1000, xres=128, yres=128, outfile=name + ".pfm")
        assert 'PixelFilter "box"' in t
        t = t.replace('PixelFilter "box"', 'PixelFilter "gaussian" "float xwidth" [9] "float ywidth" [9]')
        scene = tmp_path / (name + ".pbrt"); scene.write_text(t)
        env = dict(os.environ, LD_LIBRARY_PATH=str(mock)); env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
        out = subprocess.run([BIN, "--quiet", "--ncores", cores, str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-1500:]
        imgs.append((tmp_path / (name + ".pfm")).read_bytes())
    assert imgs[0] == imgs[1] and len(imgs[0]) > 128 * 128 * 12


def test_deferred_secondary_rays_equal_the_recursive_evaluation(mock, tmp_path, pkg):
    """A/B of the two ways the drop-in evaluates what lies behind a specular bounce.  PV_DEFER_SECONDARY=0: the ray's volume term is a
    blocking batched call inside the reference's recursion, its transmittance multiplies the throughput of the next hit at once, and
    T * Ls + Lv travels back up through SpecularReflect / SpecularTransmit.  Default: the ray is deferred to the group's single
    pv_gather_indexed call and the transmittances are folded into the queued weights afterwards, chain by chain.  Same random
    numbers, same device answers (the double's volume term is a function of the ray's stream index with T = 0.75) -- the frames
    must agree to float rounding, with final gathering on and off."""
    import numpy as np
    from cs348b_pbrt_b200 import scenes
    from test_dropin_render import read_pfm
    for fg in (True, False):
        t = scenes.cornell_surf_pbrt(nphotons=1000, caustic=300, indirect=500, finalgather=fg, fgsamples=4, xres=96, yres=96, outfile="ab.pfm")
        scene = tmp_path / "ab.pbrt"; scene.write_text(t)
        imgs, calls_n = [], []
        for defer in ("0", "1"):
            log = tmp_path / ("ab%s%d.log" % (defer, fg))
            env = dict(os.environ, LD_LIBRARY_PATH=str(mock), MOCK_PV_LOG=str(log), PV_DEFER_SECONDARY=defer)
            env.pop("PV_DEVICES", None); env.pop("PV_DEVICE", None)
            out = subprocess.run([BIN, "--quiet", "--ncores", "2", str(scene)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
            assert out.returncode == 0, out.stderr[-1500:]
            imgs.append(read_pfm(str(tmp_path / "ab.pfm")))
            calls_n.append(len([c for c in calls(log.read_text().splitlines(), "gather") if c.get("indexed")]))
        assert calls_n[0] > 100 and calls_n[1] == 1                                # thousands of blocking calls vs one per group of tasks
        a, b = imgs
        assert a.mean() > 0.1 and np.abs(a - b).max() <= 2e-5 * max(1.0, float(np.abs(a).max()))
        assert (np.abs(a - b) / (np.abs(a) + 1e-6)).max() < 1e-5
