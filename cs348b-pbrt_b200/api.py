"""Host-side mirror of the reference's PhotonShooter + PhotonVolumeIntegrator over the C ABI.

Names follow the reference: `Preprocess` (core/photonshooter.cpp:457-526), `Li`
and `Transmittance` (integrators/photonvolume.cpp:15-30,112-222), `LPhoton`
(:65-108), `Lookup` (core/kdtree.h:150-183), `Intersect`/`IntersectP`
(accelerators/bvh.cpp:585-685).  Everything forwards to csrc/libpv.so; nothing
is computed here.
"""
import ctypes as C
import os
import numpy as np
from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PVError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pv error %d: %s" % (code, msg))
        self.code = code


def library_path():
    """csrc/libpv.so; PV_LIBPV names another build of the same library (kernel tuning variants, see csrc/variants.sh)."""
    return os.environ.get("PV_LIBPV") or os.path.join(_HERE, "csrc", "libpv.so")


def load_library():
    """Load csrc/libpv.so.  Raises if it has not been built -- there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise PVError(A.PV_ESTATE, "CUDA library %s is missing: run `python __graft_entry__.py` (build()) first" % path)
    L = C.CDLL(path)
    L.pv_last_error.restype = C.c_char_p
    L.pv_last_error.argtypes = [C.c_void_p]
    L.pv_stream.restype = C.c_void_p
    L.pv_stream.argtypes = [C.c_void_p]
    L.pv_destroy.argtypes = [C.c_void_p]
    L.pv_destroy.restype = None
    _LIB = L
    return L


def _vp(a):
    """void* of a numpy array (host) or a torch tensor / int (device)."""
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class PhotonVolume:
    """One pv_ctx: scene + photon map + the integrator parameters."""

    def __init__(self, device=0, stepsize=1.0, nused=250, maxdist=0.1, seed=0):
        # defaults == CreatePhotonVolumeIntegrator (integrators/photonvolume.cpp:224-229)
        self.lib = load_library()
        self.ctx = C.c_void_p(None)
        rc = self.lib.pv_create(C.byref(self.ctx), C.c_int(device))
        if rc != 0:
            raise PVError(rc, (self.lib.pv_last_error(None) or b"").decode())
        self.stepsize, self.nused, self.maxdist, self.seed = float(stepsize), int(nused), float(maxdist), int(seed)
        self.scene = None

    def close(self):
        if self.ctx:
            self.lib.pv_destroy(self.ctx)
            self.ctx = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise PVError(rc, (self.lib.pv_last_error(self.ctx) or b"").decode())

    # ---- scene ---------------------------------------------------------
    def set_scene(self, scene):
        self.scene = scene
        d = scene.desc()
        self._chk(self.lib.pv_set_scene(self.ctx, C.byref(d)))

    # ---- photon map ----------------------------------------------------
    def set_photons(self, pos, wi, alpha):
        pos, wi, alpha = _f32(pos), _f32(wi), _f32(alpha)
        n = pos.size // 3
        self._chk(self.lib.pv_set_photons(self.ctx, _vp(pos), _vp(wi), _vp(alpha), C.c_uint64(n)))

    def set_photons_dev(self, pos, wi, alpha, n):
        self._chk(self.lib.pv_set_photons_dev(self.ctx, _vp(pos), _vp(wi), _vp(alpha), C.c_uint64(n)))

    def photon_count(self):
        n = C.c_uint64(0)
        self._chk(self.lib.pv_photon_count(self.ctx, C.byref(n)))
        return n.value

    def get_photons(self):
        n = self.photon_count()
        pos = np.zeros((n, 3), np.float32); wi = np.zeros((n, 3), np.float32)
        alpha = np.zeros((n, A.NSPEC), np.float32); ids = np.zeros(n, np.uint64)
        got = C.c_uint64(0)
        self._chk(self.lib.pv_get_photons(self.ctx, _vp(pos), _vp(wi), _vp(alpha), _vp(ids), C.c_uint64(n), C.byref(got)))
        return pos, wi, alpha, ids

    def get_photons_dev(self, pos, wi, alpha, ids, capacity):
        got = C.c_uint64(0)
        self._chk(self.lib.pv_get_photons_dev(self.ctx, _vp(pos), _vp(wi), _vp(alpha), _vp(ids), C.c_uint64(capacity), C.byref(got)))
        return got.value

    def build(self, maxdist=None, nused=None):
        maxdist = self.maxdist if maxdist is None else float(maxdist)
        nused = self.nused if nused is None else int(nused)
        self._chk(self.lib.pv_build(self.ctx, C.c_float(maxdist), C.c_uint32(nused)))

    # ---- PhotonShooter::Preprocess ---------------------------------------
    def Preprocess(self, n_volume_wanted, stepsize=0.1, max_photon_depth=5, rank=0, world=1, max_paths=0,
                   build=True, integrator_stepsize=None):
        prm = A.ShootParams(float(stepsize), float(self.stepsize if integrator_stepsize is None else integrator_stepsize),
                            int(max_photon_depth), self.seed, rank, world, max_paths, 0.0)
        st = A.ShootStats()
        self._chk(self.lib.pv_shoot(self.ctx, C.c_uint64(n_volume_wanted), C.byref(prm), C.byref(st)))
        if build and world == 1:
            self.build()
        return st

    # ---- PhotonShooter::Preprocess with the surface maps on (core/photonshooter.cpp:147-189, 457-526) ----
    def PreprocessMaps(self, n_volume_wanted, n_caustic_wanted, n_indirect_wanted, final_gather, stepsize=0.1, max_photon_depth=5,
                       max_paths=0, integrator_stepsize=None):
        prm = A.ShootParams(float(stepsize), float(self.stepsize if integrator_stepsize is None else integrator_stepsize),
                            int(max_photon_depth), self.seed, 0, 1, max_paths, 0.0)
        mp = A.MapsParams(int(n_volume_wanted), int(n_caustic_wanted), int(n_indirect_wanted), 1 if final_gather else 0)
        st = A.MapsStats()
        self._chk(self.lib.pv_shoot_maps(self.ctx, C.byref(mp), C.byref(prm), C.byref(st)))
        return st

    def PreprocessMapsRanks(self, n_volume_wanted, n_caustic_wanted, n_indirect_wanted, final_gather, rank, world, allreduce, stepsize=0.1,
                            max_photon_depth=5, max_paths=0, integrator_stepsize=None):
        """pv_shoot_maps_ranks: the all-maps pass sharded by 4096-path blocks.  `allreduce(np.ndarray[uint32])` sums the array in
        place over all ranks (see multigpu.allreduce_counts)."""
        prm = A.ShootParams(float(stepsize), float(self.stepsize if integrator_stepsize is None else integrator_stepsize),
                            int(max_photon_depth), self.seed, int(rank), int(world), max_paths, 0.0)
        mp = A.MapsParams(int(n_volume_wanted), int(n_caustic_wanted), int(n_indirect_wanted), 1 if final_gather else 0)
        st = A.MapsStats()

        def _cb(data, n, user):
            try:
                allreduce(np.ctypeslib.as_array(data, shape=(int(n),)))
                return 0
            except Exception:                      # never unwind through the C frames
                import traceback; traceback.print_exc()
                return 1
        cb = A.ALLREDUCE_U32_FN(_cb)
        self._chk(self.lib.pv_shoot_maps_ranks(self.ctx, C.byref(mp), C.byref(prm), cb, None, C.byref(st)))
        return st

    def get_map_photons(self, which, capacity=None):
        n = C.c_uint64(0)
        self._chk(self.lib.pv_get_map_photons(self.ctx, C.c_int(which), None, None, None, None, C.c_uint64(1 << 62), C.byref(n)))
        n = n.value if capacity is None else min(n.value, capacity)
        pos = np.zeros((n, 3), np.float32); wi = np.zeros((n, 3), np.float32)
        alpha = np.zeros((n, A.NSPEC), np.float32); ids = np.zeros(n, np.uint64)
        got = C.c_uint64(0)
        self._chk(self.lib.pv_get_map_photons(self.ctx, C.c_int(which), _vp(pos), _vp(wi), _vp(alpha), _vp(ids), C.c_uint64(n), C.byref(got)))
        return pos, wi, alpha, ids

    def set_map_photons(self, which, pos, wi, alpha):
        pos, wi, alpha = _f32(pos), _f32(wi), _f32(alpha)
        self._chk(self.lib.pv_set_map_photons(self.ctx, C.c_int(which), _vp(pos), _vp(wi), _vp(alpha), C.c_uint64(pos.size // 3)))

    def RadiancePhotons(self, n_lookup, max_dist2, path_counts=None):
        """ComputeRadianceTask (core/photonshooter.cpp:359-395): Lo of every radiance photon.  path_counts = (nDirectPaths,
        nIndirectPaths, nCausticPaths) or None for those of the last PreprocessMaps."""
        n = C.c_uint64(0)
        self._chk(self.lib.pv_get_map_photons(self.ctx, C.c_int(A.MAP_RADIANCE), None, None, None, None, C.c_uint64(1 << 62), C.byref(n)))
        Lo = np.zeros((n.value, A.NSPEC), np.float32)
        pc = (C.c_uint64 * 3)(*[int(c) for c in path_counts]) if path_counts is not None else None
        got = C.c_uint64(0)
        self._chk(self.lib.pv_radiance_photons(self.ctx, C.c_uint32(n_lookup), C.c_float(max_dist2), pc, _vp(Lo), C.c_uint64(n.value), C.byref(got)))
        return Lo

    def set_radiance_lo(self, Lo):
        Lo = _f32(Lo).reshape(-1, A.NSPEC)
        self._chk(self.lib.pv_set_radiance_lo(self.ctx, _vp(Lo), C.c_uint64(len(Lo))))

    def select_map(self, which, maxdist, nused):
        """Build the lookup grid over a photon class (A.MAP_*); build() puts the volume map back."""
        self._chk(self.lib.pv_select_map(self.ctx, C.c_int(which), C.c_float(maxdist), C.c_uint32(nused)))

    def SurfaceLPhoton(self, pts, nf, n_lookup, max_dist2, n_paths):
        """PhotonIntegrator's LPhoton, diffuse branch (integrators/photonmap.cpp:62-108) on the selected surface map -> (Lr, Lt)."""
        pts = _f32(pts).reshape(-1, 3); nf = _f32(nf).reshape(-1, 3); n = len(pts)
        Lr = np.zeros((n, A.NSPEC), np.float32); Lt = np.zeros((n, A.NSPEC), np.float32)
        self._chk(self.lib.pv_surface_lphoton(self.ctx, _vp(pts), _vp(nf), C.c_uint64(n), C.c_uint32(n_lookup), C.c_float(max_dist2),
                                              C.c_uint64(int(n_paths)), _vp(Lr), _vp(Lt)))
        return Lr, Lt

    def RadianceNearest(self, pts, normals, want_Lo=True):
        """RadiancePhotonProcess lookup of final gathering (integrators/photonmap.cpp:238-243) -> (index, Lo)."""
        pts = _f32(pts).reshape(-1, 3); normals = _f32(normals).reshape(-1, 3); n = len(pts)
        idx = np.zeros(n, np.uint32); Lo = np.zeros((n, A.NSPEC), np.float32) if want_Lo else None
        self._chk(self.lib.pv_radiance_nearest(self.ctx, _vp(pts), _vp(normals), C.c_uint64(n), _vp(idx), _vp(Lo)))
        return idx, Lo

    def FinalGather(self, rays, step=None, index_base=0):
        """One batch of final-gather rays (integrators/photonmap.cpp:231-243): trace, nearest facing radiance photon at the hit, its
        Lo times the transmittance along the ray -> (Lindir, radiance photon index)."""
        rays = np.ascontiguousarray(rays); n = len(rays)
        L = np.zeros((n, A.NSPEC), np.float32); idx = np.zeros(n, np.uint32)
        step = 4.0 * self.stepsize if step is None else float(step)
        self._chk(self.lib.pv_final_gather(self.ctx, _vp(rays), C.c_uint64(n), C.c_float(step), C.c_uint64(self.seed), C.c_uint64(index_base),
                                           _vp(L), _vp(idx)))
        return L, idx

    # ---- KdTree::Lookup --------------------------------------------------
    def Lookup(self, pts, k=None, r2=None):
        pts = _f32(pts).reshape(-1, 3); n = len(pts)
        k = self.nused if k is None else int(k)
        r2 = self.maxdist * self.maxdist if r2 is None else float(r2)
        idx = np.zeros((n, k), np.uint32); d2 = np.zeros((n, k), np.float32); nf = np.zeros(n, np.uint32)
        self._chk(self.lib.pv_knn(self.ctx, _vp(pts), C.c_uint64(n), C.c_uint32(k), C.c_float(r2), _vp(idx), _vp(d2), _vp(nf)))
        return nf, idx, d2

    def LPhoton(self, pts, w):
        pts = _f32(pts).reshape(-1, 3); w = _f32(w).reshape(-1, 3); n = len(pts)
        L = np.zeros((n, A.NSPEC), np.float32)
        self._chk(self.lib.pv_lphoton(self.ctx, _vp(pts), _vp(w), C.c_uint64(n), C.c_uint32(self.nused), C.c_float(self.maxdist), _vp(L)))
        return L

    # ---- BVHAccel::BVHAccel on the device (accelerators/bvh.cpp:196-577) --------
    def build_bvh(self, prim_bounds, max_prims_in_node=4):
        """LBVH over primitive bounds [n, 6] (pMin, pMax) -> (nodes: uint8 [32 * n_nodes] in the LinearBVHNode layout,
        prim_order: uint32 [n], kernel time in ms).  max_prims_in_node as the reference's "maxnodeprims" (default 4, bvh.cpp:694)."""
        pb = _f32(prim_bounds).reshape(-1, 6); n = len(pb)
        nodes = np.zeros(32 * max(2 * n - 1, 1), np.uint8); order = np.zeros(n, np.uint32)
        n_nodes = C.c_uint32(0); ms = C.c_float(0.0)
        self._chk(self.lib.pv_build_bvh(self.ctx, _vp(pb), C.c_uint32(n), C.c_uint32(max_prims_in_node), _vp(nodes),
                                        C.c_uint32(max(2 * n - 1, 1)), C.byref(n_nodes), _vp(order), C.byref(ms)))
        return nodes[:32 * n_nodes.value].copy(), order, float(ms.value)

    # ---- Scene::Intersect / IntersectP -----------------------------------
    def Intersect(self, rays):
        rays = np.ascontiguousarray(rays); n = len(rays)
        prim = np.zeros(n, np.uint32); t = np.zeros(n, np.float32)
        self._chk(self.lib.pv_intersect(self.ctx, _vp(rays), C.c_uint64(n), _vp(prim), _vp(t)))
        return prim, t

    def IntersectP(self, rays):
        rays = np.ascontiguousarray(rays); n = len(rays)
        hit = np.zeros(n, np.uint8)
        self._chk(self.lib.pv_occluded(self.ctx, _vp(rays), C.c_uint64(n), _vp(hit)))
        return hit

    # ---- PhotonVolumeIntegrator::Transmittance / Li ------------------------
    def Transmittance(self, rays, offset_u, step=None):
        rays = np.ascontiguousarray(rays); n = len(rays)
        u = _f32(offset_u)
        T = np.zeros((n, A.NSPEC), np.float32)
        step = 4.0 * self.stepsize if step is None else float(step)   # sample == NULL branch, photonvolume.cpp:24-27
        self._chk(self.lib.pv_transmittance(self.ctx, _vp(rays), C.c_uint64(n), C.c_float(step), _vp(u), _vp(T)))
        return T

    def gather_params(self, ray_index_base=0, flags=0):
        return A.GatherParams(self.stepsize, self.nused, self.maxdist, self.seed, ray_index_base, flags)

    def Li(self, rays, ray_index_base=0, flags=0):
        """Host buffers in, host buffers out (copies inside)."""
        rays = np.ascontiguousarray(rays); n = len(rays)
        L = np.zeros((n, A.NSPEC), np.float32); T = np.zeros((n, A.NSPEC), np.float32)
        prm = self.gather_params(ray_index_base, flags)
        self._chk(self.lib.pv_gather(self.ctx, _vp(rays), C.c_uint64(n), C.byref(prm), _vp(L), _vp(T)))
        return L, T

    def LiIndexed(self, rays, ray_index, flags=0, integrator=None):
        """pv_gather_indexed / pv_volume_li_indexed: ray i draws from the stream ray_index[i]."""
        rays = np.ascontiguousarray(rays); n = len(rays)
        idx = np.ascontiguousarray(ray_index, dtype=np.uint64)
        assert len(idx) == n
        L = np.zeros((n, A.NSPEC), np.float32); T = np.zeros((n, A.NSPEC), np.float32)
        prm = self.gather_params(0, flags)
        if integrator is None:
            self._chk(self.lib.pv_gather_indexed(self.ctx, _vp(rays), _vp(idx), C.c_uint64(n), C.byref(prm), _vp(L), _vp(T)))
        else:
            kind = {"single": A.VOLINT_SINGLE, "emission": A.VOLINT_EMISSION}[integrator]
            self._chk(self.lib.pv_volume_li_indexed(self.ctx, C.c_int(kind), _vp(rays), _vp(idx), C.c_uint64(n), C.byref(prm), _vp(L), _vp(T)))
        return L, T

    def VolumeLi(self, integrator, rays, ray_index_base=0, flags=0):
        """SingleScatteringIntegrator::Li / EmissionIntegrator::Li (integrators/single.cpp:66-138, emission.cpp:63-106):
        integrator = "single" | "emission"; uses this object's stepsize and seed, no photon map."""
        kind = {"single": A.VOLINT_SINGLE, "emission": A.VOLINT_EMISSION}[integrator]
        rays = np.ascontiguousarray(rays); n = len(rays)
        L = np.zeros((n, A.NSPEC), np.float32); T = np.zeros((n, A.NSPEC), np.float32)
        prm = self.gather_params(ray_index_base, flags)          # flags: A.VOLINT_WARP_PER_RAY / A.VOLINT_THREAD_PER_RAY
        self._chk(self.lib.pv_volume_li(self.ctx, C.c_int(kind), _vp(rays), C.c_uint64(n), C.byref(prm), _vp(L), _vp(T)))
        return L, T

    def Li_into(self, rays, n, L, T, ray_index_base=0, flags=0):
        """Host pointers (e.g. pinned torch tensors / numpy) -- no allocation."""
        prm = self.gather_params(ray_index_base, flags)
        self._chk(self.lib.pv_gather(self.ctx, _vp(rays), C.c_uint64(n), C.byref(prm), _vp(L), _vp(T)))

    def Li_dev(self, rays, n, L, T, ray_index_base=0, flags=0):
        """Device pointers (torch CUDA tensors or ints)."""
        prm = self.gather_params(ray_index_base, flags)
        self._chk(self.lib.pv_gather_dev(self.ctx, _vp(rays), C.c_uint64(n), C.byref(prm), _vp(L), _vp(T)))

    def gather_stats(self, reset=False):
        st = A.GatherStats()
        self._chk(self.lib.pv_gather_stats_get(self.ctx, C.byref(st), C.c_int(1 if reset else 0)))
        return st

    def last_kernel_ms(self):
        ms = C.c_float(0)
        self._chk(self.lib.pv_last_kernel_ms(self.ctx, C.byref(ms)))
        return ms.value

    def last_march_ms(self):
        ms = C.c_float(0)
        self._chk(self.lib.pv_last_march_ms(self.ctx, C.byref(ms)))
        return ms.value

    def last_phase_ms(self):
        """Cell-batched schedule: device ms of (step sort, cellgather_kernel, overflow pass, recurrence) of the last Li."""
        ms = (C.c_float * 4)()
        self._chk(self.lib.pv_last_phase_ms(self.ctx, ms))
        return [float(v) for v in ms]

    # ---- NCCL behind the C ABI (csrc/pv_comm.cu) ---------------------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes for pv_comm_init, made by rank 0 and handed to the other ranks by the caller's own means."""
        lib = load_library()
        buf = (C.c_uint8 * 128)()
        rc = lib.pv_comm_unique_id(buf)
        if rc != 0:
            raise PVError(rc, (lib.pv_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._chk(self.lib.pv_comm_init(self.ctx, buf, C.c_int(rank), C.c_int(world)))

    def comm_destroy(self):
        self._chk(self.lib.pv_comm_destroy(self.ctx))

    def allgather_photons(self, renumber=False):
        """Every rank ends with the union of all ranks' photons ordered by id; returns the device ms of the collective."""
        ms = C.c_float(0)
        self._chk(self.lib.pv_allgather_photons(self.ctx, C.c_int(1 if renumber else 0), C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_uint64(0)
        self._chk(self.lib.pv_launch_count(self.ctx, C.byref(n)))
        return n.value

    def stream(self):
        return self.lib.pv_stream(self.ctx)
