"""CPU: the C-ABI library builds, loads and exports every symbol include/pv.h declares; struct layouts match."""
import ctypes as C
import os
import re
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_listed(pkg):
    assert declared_symbols() == sorted(pkg._abi.EXPORTS)


def test_library_exports_every_declared_symbol(pkg):
    path = pkg.library_path()
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.dirname(path), "-j8"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_sizes_match_the_c_compiler(pkg, tmp_path):
    A = pkg._abi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "pv.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(pv_bvh_node),sizeof(pv_ray),sizeof(pv_light),sizeof(pv_medium),sizeof(pv_material),'
                   'sizeof(pv_scene_desc),sizeof(pv_gather_params),sizeof(pv_shoot_params),sizeof(pv_shoot_stats),'
                   'sizeof(pv_gather_stats),sizeof(pv_maps_params),sizeof(pv_maps_stats),sizeof(pv_sphere));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mine = [C.sizeof(t) for t in (A.BvhNode, A.Ray, A.Light, A.Medium, A.Material, A.SceneDesc, A.GatherParams,
                                  A.ShootParams, A.ShootStats, A.GatherStats, A.MapsParams, A.MapsStats, A.Sphere)]
    assert sizes == mine
    assert sizes[0] == 32 and sizes[1] == 40


def test_no_gpu_means_a_loud_error_not_a_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.PVError) as e:
        pkg.PhotonVolume(device=0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_the_oracle():
    """The product path must not reach into oracle/ (it is test infrastructure)."""
    pkgdir = os.path.join(ROOT, "cs348b-pbrt_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("pv_oracle", "oracle_lib", "libpv_oracle", "import oracle", "from oracle", "oracle/_ref"):
                    assert needle not in text, (f, needle)
