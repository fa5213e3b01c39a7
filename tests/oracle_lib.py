"""ctypes binding of oracle/_ref/libpv_oracle.so -- the CPU checker.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
A = pkg._abi

_LIB = None
PHILOX, MT = 0, 1


class Photons(C.Structure):
    _fields_ = [("n", C.c_uint64), ("pos", C.POINTER(C.c_float)), ("wi", C.POINTER(C.c_float)),
                ("alpha", C.POINTER(C.c_float)), ("ids", C.POINTER(C.c_uint64)), ("nshot", C.c_uint64),
                ("blocks", C.c_uint64), ("nodes_visited", C.c_uint64), ("tri_tests", C.c_uint64),
                ("density_samples", C.c_uint64), ("segments", C.c_uint64)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_ref", "libpv_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.pvo_kdtree_build.restype = C.c_void_p
        L.pvo_kdtree_build.argtypes = [C.c_void_p, C.c_uint64]
        L.pvo_kdtree_free.argtypes = [C.c_void_p]
        L.pvo_spectrum_y.restype = C.c_float
        L.pvo_van_der_corput.restype = C.c_float
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class KdTree:
    def __init__(self, pos):
        self.pos = f32(pos).reshape(-1, 3)
        self.h = lib().pvo_kdtree_build(_p(self.pos), C.c_uint64(len(self.pos)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().pvo_kdtree_free(C.c_void_p(self.h)); self.h = None

    def knn(self, pts, k, r2):
        pts = f32(pts).reshape(-1, 3); n = len(pts)
        idx = np.zeros((n, k), np.uint32); d2 = np.zeros((n, k), np.float32); nf = np.zeros(n, np.uint32)
        ties = C.c_uint64(0)
        lib().pvo_knn(C.c_void_p(self.h), _p(pts), C.c_uint64(n), C.c_uint32(k), C.c_float(r2), _p(idx), _p(d2), _p(nf), C.byref(ties))
        return nf, idx, d2, ties.value


def knn_brute(pos, pts, k, r2):
    pos = f32(pos).reshape(-1, 3); pts = f32(pts).reshape(-1, 3); n = len(pts)
    idx = np.zeros((n, k), np.uint32); d2 = np.zeros((n, k), np.float32); nf = np.zeros(n, np.uint32)
    lib().pvo_knn_brute(_p(pos), C.c_uint64(len(pos)), _p(pts), C.c_uint64(n), C.c_uint32(k), C.c_float(r2), _p(idx), _p(d2), _p(nf))
    return nf, idx, d2


def lphoton(scene, tree, wi, alpha, pts, w, nused, maxdist):
    pts = f32(pts).reshape(-1, 3); w = f32(w).reshape(-1, 3); n = len(pts)
    wi = f32(wi); alpha = f32(alpha)
    L = np.zeros((n, A.NSPEC), np.float32)
    d = scene.desc()
    lib().pvo_lphoton(C.byref(d), C.c_void_p(tree.h), _p(wi), _p(alpha), _p(pts), _p(w), C.c_uint64(n), C.c_uint32(nused), C.c_float(maxdist), _p(L))
    return L


def intersect(scene, rays):
    n = len(rays); rays = np.ascontiguousarray(rays)
    prim = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); occ = np.zeros(n, np.uint8)
    d = scene.desc()
    lib().pvo_intersect(C.byref(d), _p(rays), C.c_uint64(n), _p(prim), _p(t))
    lib().pvo_occluded(C.byref(d), _p(rays), C.c_uint64(n), _p(occ))
    return prim, t, occ


def transmittance(scene, rays, step, offset_u):
    n = len(rays); rays = np.ascontiguousarray(rays); u = f32(offset_u)
    T = np.zeros((n, A.NSPEC), np.float32)
    d = scene.desc()
    lib().pvo_transmittance(C.byref(d), _p(rays), C.c_uint64(n), C.c_float(step), _p(u), _p(T))
    return T


def gather(scene, tree, wi, alpha, rays, stepsize, nused, maxdist, seed=0, ray_index_base=0, flags=0,
           rng_mode=PHILOX, mt_seed=0, nthreads=1):
    n = len(rays); rays = np.ascontiguousarray(rays)
    wi = f32(wi); alpha = f32(alpha)
    L = np.zeros((n, A.NSPEC), np.float32); T = np.zeros((n, A.NSPEC), np.float32)
    prm = A.GatherParams(stepsize, nused, maxdist, seed, ray_index_base, flags)
    st = A.GatherStats()
    d = scene.desc()
    lib().pvo_gather(C.byref(d), C.c_void_p(tree.h if tree is not None else None), _p(wi), _p(alpha), _p(rays), C.c_uint64(n),
                     C.byref(prm), C.c_int(rng_mode), C.c_uint32(mt_seed), C.c_int(nthreads), _p(L), _p(T), C.byref(st))
    return L, T, st


SINGLE, EMISSION = 0, 1


class AreaLight(C.Structure):
    _fields_ = [("slot", C.c_uint32), ("n_tris", C.c_uint32), ("flags", C.c_uint32), ("pad", C.c_uint32),
                ("tri", C.POINTER(C.c_float)), ("Lemit", C.c_float * 30)]


class area_lights:
    """DiffuseAreaLights for the oracle: `with area_lights(path_of_PVAREA01_file): ...` (written by ref_harness --export-area-lights
    next to a scene whose area-light slots hold placeholders).  pv_light has no area-light fields yet."""

    def __init__(self, path):
        buf = open(path, "rb").read()
        assert buf[:8] == b"PVAREA01"
        n = int(np.frombuffer(buf, np.uint64, 1, 8)[0])
        off = 16
        self.keep = []
        self.arr = (AreaLight * n)()
        for i in range(n):
            slot, nt, flags = np.frombuffer(buf, np.uint32, 3, off); off += 12
            lem = np.frombuffer(buf, np.float32, 30, off).copy(); off += 120
            tri = np.frombuffer(buf, np.float32, 9 * int(nt), off).copy(); off += 36 * int(nt)
            self.keep.append(tri)
            a = self.arr[i]
            a.slot, a.n_tris, a.flags, a.pad = int(slot), int(nt), int(flags), 0
            a.tri = tri.ctypes.data_as(C.POINTER(C.c_float))
            for b in range(30):
                a.Lemit[b] = float(lem[b])

    def __enter__(self):
        lib().pvo_set_area_lights(self.arr, C.c_uint32(len(self.arr)))
        return self

    def __exit__(self, *exc):
        lib().pvo_set_area_lights(None, C.c_uint32(0))
        return False


class more_media:
    """AggregateVolume for the oracle (core/volume.cpp:178-261): `with more_media(scene_b, scene_c): ...` makes every oracle call
    inside see the medium of the scene it is given PLUS the media of these scenes as one aggregate.  pv_scene_desc carries one
    medium (the device path has no aggregate yet), so the further regions travel on the side."""

    def __init__(self, *scenes):
        self.scenes = scenes

    def __enter__(self):
        descs = [s.desc() for s in self.scenes]               # keeps the density pointers alive through the scenes
        self.arr = (A.Medium * len(descs))(*[d.medium.contents for d in descs])
        lib().pvo_set_more_media(self.arr, C.c_uint32(len(descs)))
        return self

    def __exit__(self, *exc):
        lib().pvo_set_more_media(None, C.c_uint32(0))
        return False


def volume_li(scene, rays, stepsize, kind, seed=0, ray_index_base=0, rng_mode=PHILOX, mt_seed=0):
    """SingleScatteringIntegrator::Li / EmissionIntegrator::Li (integrators/single.cpp:66-138, emission.cpp:63-106)."""
    n = len(rays); rays = np.ascontiguousarray(rays)
    L = np.zeros((n, A.NSPEC), np.float32); T = np.zeros((n, A.NSPEC), np.float32)
    prm = A.GatherParams(stepsize, 0, 0.0, seed, ray_index_base, 0)
    st = A.GatherStats()
    d = scene.desc()
    rc = lib().pvo_volume_li(C.byref(d), _p(rays), C.c_uint64(n), C.byref(prm), C.c_int(kind), C.c_int(rng_mode), C.c_uint32(mt_seed),
                             _p(L), _p(T), C.byref(st))
    assert rc == 0
    return L, T, st


def shoot(scene, n_wanted, stepsize, integrator_stepsize, max_photon_depth=5, seed=0, rng_mode=PHILOX, nthreads=1,
          max_paths=0):
    prm = A.ShootParams(stepsize, integrator_stepsize, max_photon_depth, seed, 0, 1, max_paths, 0.0)
    out = Photons()
    d = scene.desc()
    rc = lib().pvo_shoot(C.byref(d), C.c_uint64(n_wanted), C.byref(prm), C.c_int(rng_mode), C.c_int(nthreads), C.byref(out))
    n = out.n
    res = dict(rc=rc, n=n, nshot=out.nshot, blocks=out.blocks, nodes_visited=out.nodes_visited, tri_tests=out.tri_tests,
               density_samples=out.density_samples, segments=out.segments)
    if n:
        res["pos"] = np.ctypeslib.as_array(out.pos, shape=(n, 3)).copy()
        res["wi"] = np.ctypeslib.as_array(out.wi, shape=(n, 3)).copy()
        res["alpha"] = np.ctypeslib.as_array(out.alpha, shape=(n, A.NSPEC)).copy()
        res["ids"] = np.ctypeslib.as_array(out.ids, shape=(n,)).copy()
    else:
        res["pos"] = np.zeros((0, 3), np.float32); res["wi"] = np.zeros((0, 3), np.float32)
        res["alpha"] = np.zeros((0, A.NSPEC), np.float32); res["ids"] = np.zeros(0, np.uint64)
    lib().pvo_photons_free(C.byref(out))
    return res


class Maps(C.Structure):
    _fields_ = [("cls", Photons * 5), ("nshot", C.c_uint64), ("blocks", C.c_uint64), ("n_caustic_paths", C.c_uint64),
                ("n_indirect_paths", C.c_uint64), ("n_direct_paths", C.c_uint64), ("n_volume_paths", C.c_uint64)]


MAP_NAMES = ("volume", "caustic", "indirect", "direct", "radiance")


def shoot_maps(scene, n_volume, n_caustic, n_indirect, final_gather, stepsize, integrator_stepsize, max_photon_depth=5, seed=0,
               rng_mode=PHILOX, max_paths=0):
    """pvo_shoot_maps: all photon maps of one PhotonShootingTask (core/photonshooter.cpp:232-357)."""
    prm = A.ShootParams(stepsize, integrator_stepsize, max_photon_depth, seed, 0, 1, max_paths, 0.0)
    out = Maps()
    d = scene.desc()
    rc = lib().pvo_shoot_maps(C.byref(d), C.c_uint64(n_volume), C.c_uint64(n_caustic), C.c_uint64(n_indirect), C.c_int(int(final_gather)),
                              C.byref(prm), C.c_int(rng_mode), C.byref(out))
    res = dict(rc=rc, nshot=out.nshot, blocks=out.blocks, caustic_paths=out.n_caustic_paths, indirect_paths=out.n_indirect_paths,
               direct_paths=out.n_direct_paths, volume_paths=out.n_volume_paths)
    for k, name in enumerate(MAP_NAMES):
        c = out.cls[k]; n = c.n
        if n:
            res[name] = dict(pos=np.ctypeslib.as_array(c.pos, shape=(n, 3)).copy(), wi=np.ctypeslib.as_array(c.wi, shape=(n, 3)).copy(),
                             alpha=np.ctypeslib.as_array(c.alpha, shape=(n, A.NSPEC)).copy(), ids=np.ctypeslib.as_array(c.ids, shape=(n,)).copy())
        else:
            res[name] = dict(pos=np.zeros((0, 3), np.float32), wi=np.zeros((0, 3), np.float32), alpha=np.zeros((0, A.NSPEC), np.float32),
                             ids=np.zeros(0, np.uint64))
    lib().pvo_maps_free(C.byref(out))
    return res


def radiance(maps, counts, rp_pos, rp_n, rho_r, n_lookup, max_dist2):
    """pvo_radiance: maps = [(pos, wi, alpha) or None] in the order direct, indirect, caustic; counts = path counts."""
    trees, handles, wis, alphas = [], (C.c_void_p * 3)(), (C.c_void_p * 3)(), (C.c_void_p * 3)()
    keep = []
    for k, m in enumerate(maps):
        if m is None or len(m[0]) == 0:
            handles[k] = None; wis[k] = None; alphas[k] = None
            continue
        t = KdTree(m[0]); w = f32(m[1]); a = f32(m[2])
        trees.append(t); keep += [w, a]
        handles[k] = t.h; wis[k] = w.ctypes.data; alphas[k] = a.ctypes.data
    cnt = (C.c_uint64 * 3)(*[int(c) for c in counts])
    rp_pos = f32(rp_pos).reshape(-1, 3); rp_n = f32(rp_n).reshape(-1, 3); rho_r = f32(rho_r).reshape(-1, A.NSPEC)
    n = len(rp_pos)
    Lo = np.zeros((n, A.NSPEC), np.float32)
    lib().pvo_radiance(handles, wis, alphas, cnt, _p(rp_pos), _p(rp_n), _p(rho_r), C.c_uint64(n), C.c_uint32(n_lookup), C.c_float(max_dist2), _p(Lo))
    return Lo


def surface_lphoton(pos, wi, alpha, pts, nf, n_lookup, max_dist2, n_paths):
    """pvo_surface_lphoton: PhotonIntegrator's LPhoton, diffuse branch -> (Lr, Lt)."""
    pts = f32(pts).reshape(-1, 3); nf = f32(nf).reshape(-1, 3); n = len(pts)
    t = KdTree(pos); wi = f32(wi); alpha = f32(alpha)
    Lr = np.zeros((n, A.NSPEC), np.float32); Lt = np.zeros((n, A.NSPEC), np.float32)
    lib().pvo_surface_lphoton(C.c_void_p(t.h), _p(wi), _p(alpha), _p(pts), _p(nf), C.c_uint64(n), C.c_uint32(n_lookup), C.c_float(max_dist2),
                              C.c_uint64(int(n_paths)), _p(Lr), _p(Lt))
    return Lr, Lt


def radiance_nearest(rp_pos, rp_n, pts, nrm):
    rp_pos = f32(rp_pos).reshape(-1, 3); rp_n = f32(rp_n).reshape(-1, 3); pts = f32(pts).reshape(-1, 3); nrm = f32(nrm).reshape(-1, 3)
    n = len(pts)
    idx = np.zeros(n, np.uint32); d2 = np.zeros(n, np.float32)
    lib().pvo_radiance_nearest(_p(rp_pos), _p(rp_n), C.c_uint64(len(rp_pos)), _p(pts), _p(nrm), C.c_uint64(n), _p(idx), _p(d2))
    return idx, d2


def final_gather(scene, rp_pos, rp_n, rp_Lo, rays, step, seed=0, index_base=0):
    rp_pos = f32(rp_pos).reshape(-1, 3); rp_n = f32(rp_n).reshape(-1, 3); rp_Lo = f32(rp_Lo).reshape(-1, A.NSPEC)
    rays = np.ascontiguousarray(rays); n = len(rays)
    L = np.zeros((n, A.NSPEC), np.float32); idx = np.zeros(n, np.uint32)
    d = scene.desc()
    lib().pvo_final_gather(C.byref(d), _p(rp_pos), _p(rp_n), _p(rp_Lo), C.c_uint64(len(rp_pos)), _p(rays), C.c_uint64(n), C.c_float(step),
                           C.c_uint64(seed), C.c_uint64(index_base), _p(L), _p(idx))
    return L, idx
