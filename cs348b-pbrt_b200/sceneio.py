"""Binary interchange files either side of the hot path.

PVSCN001  flattened scene exported from the reference's host objects
          (LinearBVHNode[], triangle table, materials, lights, medium [+ grid],
          optional sphere table: header word 5 = sphere count, prim_shape[] and
          pv_sphere[] appended at the end of the file; header word 6 = number of
          area-light triangles, 9 floats each after that)
PVPHOT01  photon set: n x {p.xyz, wi.xyz, alpha[30]}  (core/photonshooter.h:20-27
          minus the fork's unused lambda/intensity)
PVRAY001  rays: n x pv_ray (40 B)
PVQRY001  query points + directions: n x {p.xyz, w.xyz}
"""
import ctypes as C
import struct
import numpy as np
from . import _abi as A


class Scene:
    """Owns the numpy buffers behind a pv_scene_desc."""

    def __init__(self):
        self.nodes = np.zeros(0, dtype=np.uint8)
        self.n_nodes = 0
        self.tri_verts = np.zeros(0, dtype=np.float32)
        self.prim_material = np.zeros(0, dtype=np.uint32)
        self.materials = (A.Material * 0)()
        self.lights = (A.Light * 0)()
        self.medium = None
        self.density = None
        self.world_bound = np.zeros(6, dtype=np.float32)
        self.cie_y = np.zeros(A.NSPEC, dtype=np.float32)
        self.prim_shape = None              # uint32 per primitive: SHAPE_TRIANGLE or an index into spheres
        self.spheres = (A.Sphere * 0)()
        self.light_tris = None              # float32 [n, 9]: triangles of the area lights (PV_LIGHT_AREA), see attach_area_lights

    @property
    def n_prims(self):
        return len(self.prim_material)

    def prim_bounds(self):
        """WorldBound() of every primitive, float32 [n, 6]: a triangle's is the bound of its vertices (shapes/trianglemesh.cpp:116-121);
        a sphere's is its object-space box [-r, r]^2 x [zmin, zmax] through ObjectToWorld (shapes/sphere.cpp:53-56,
        core/transform.cpp:171-181)"""
        v = np.asarray(self.tri_verts, np.float32).reshape(-1, 3, 3)
        b = np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1).astype(np.float32)
        if self.prim_shape is not None and len(self.spheres):
            for i in np.nonzero(np.asarray(self.prim_shape) != A.SHAPE_TRIANGLE)[0]:
                sp = self.spheres[int(self.prim_shape[i])]
                m = np.array(list(sp.object_to_world), np.float32).reshape(4, 4)
                r = np.float32(sp.radius)
                c = np.array([[x, y, z, 1.0] for x in (-r, r) for y in (-r, r) for z in (sp.zmin, sp.zmax)], np.float32)
                w = c @ m.T
                w = w[:, :3] / w[:, 3:4]
                b[i, :3] = w.min(axis=0); b[i, 3:] = w.max(axis=0)
        return b

    def with_bvh(self, nodes, prim_order):
        """the same scene over another BVH: `nodes` (uint8, 32 bytes per node) whose leaves index the primitives permuted by
        prim_order (new position -> old index), as pv_build_bvh returns them"""
        import copy
        s = copy.copy(self)
        o = np.asarray(prim_order, np.int64)
        s.nodes = np.ascontiguousarray(nodes, np.uint8); s.n_nodes = len(s.nodes) // 32
        s.tri_verts = np.ascontiguousarray(np.asarray(self.tri_verts, np.float32).reshape(-1, 9)[o].reshape(-1))
        s.prim_material = np.ascontiguousarray(np.asarray(self.prim_material, np.uint32)[o])
        if self.prim_shape is not None:
            s.prim_shape = np.ascontiguousarray(np.asarray(self.prim_shape, np.uint32)[o])
        return s

    def desc(self):
        d = A.SceneDesc()
        self._nodes_buf = np.ascontiguousarray(self.nodes)
        d.nodes = C.cast(self._nodes_buf.ctypes.data, C.POINTER(A.BvhNode))
        d.n_nodes = self.n_nodes
        self.tri_verts = np.ascontiguousarray(self.tri_verts, dtype=np.float32)
        self.prim_material = np.ascontiguousarray(self.prim_material, dtype=np.uint32)
        d.tri_verts = self.tri_verts.ctypes.data_as(C.POINTER(C.c_float))
        d.prim_material = self.prim_material.ctypes.data_as(C.POINTER(C.c_uint32))
        d.n_prims = self.n_prims
        d.materials = C.cast(self.materials, C.POINTER(A.Material))
        d.n_materials = len(self.materials)
        d.lights = C.cast(self.lights, C.POINTER(A.Light))
        d.n_lights = len(self.lights)
        if self.medium is not None:
            if self.density is not None:
                self.density = np.ascontiguousarray(self.density, dtype=np.float32)
                self.medium.density = self.density.ctypes.data_as(C.POINTER(C.c_float))
            d.medium = C.pointer(self.medium)
        if len(self.spheres):
            self.prim_shape = np.ascontiguousarray(self.prim_shape, dtype=np.uint32)
            d.prim_shape = self.prim_shape.ctypes.data_as(C.POINTER(C.c_uint32))
            d.spheres = C.cast(self.spheres, C.POINTER(A.Sphere))
            d.n_spheres = len(self.spheres)
        if self.light_tris is not None and len(self.light_tris):
            self.light_tris = np.ascontiguousarray(self.light_tris, dtype=np.float32)
            d.light_tris = self.light_tris.ctypes.data_as(C.POINTER(C.c_float))
            d.n_light_tris = self.light_tris.size // 9
        for i in range(6):
            d.world_bound[i] = float(self.world_bound[i])
        for i in range(A.NSPEC):
            d.cie_y[i] = float(self.cie_y[i])
        self._desc = d
        return d


def attach_area_lights(scene, path):
    """PVAREA01 side file (oracle/ref_harness --export-area-lights: per DiffuseAreaLight its slot in the light list, Lemit and the
    triangles of its ShapeSet in refine order) -> the placeholder slots of `scene` become PV_LIGHT_AREA lights over scene.light_tris."""
    buf = open(path, "rb").read()
    assert buf[:8] == b"PVAREA01"
    n = int(np.frombuffer(buf, np.uint64, 1, 8)[0])
    off = 16
    tris = [] if scene.light_tris is None else [np.asarray(scene.light_tris, np.float32).reshape(-1, 9)]
    first = sum(len(t) for t in tris)
    for _ in range(n):
        slot, nt, flags = (int(v) for v in np.frombuffer(buf, np.uint32, 3, off)); off += 12
        lem = np.frombuffer(buf, np.float32, 30, off); off += 120
        tri = np.frombuffer(buf, np.float32, 9 * nt, off).reshape(nt, 9); off += 36 * nt
        l = scene.lights[slot]
        l.type = A.LIGHT_AREA
        bits = np.array([first, nt, flags], np.uint32).view(np.float32)        # pv_light::area shares storage with pos[3]
        for k in range(3):
            l.pos[k] = bits[k]
        for b in range(A.NSPEC):
            l.intensity[b] = float(lem[b])
        tris.append(tri); first += nt
    scene.light_tris = np.concatenate(tris) if tris else None
    return scene


def read_scene(path):
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"PVSCN001":
        raise ValueError("not a PVSCN001 file: %s" % path)
    off = 8
    hdr = struct.unpack_from("<8I", buf, off); off += 32
    n_nodes, n_prims, n_mat, n_lights, has_medium, n_spheres, n_light_tris = hdr[:7]
    s = Scene()
    s.world_bound = np.frombuffer(buf, dtype=np.float32, count=6, offset=off).copy(); off += 24
    s.cie_y = np.frombuffer(buf, dtype=np.float32, count=A.NSPEC, offset=off).copy(); off += 4 * A.NSPEC
    s.nodes = np.frombuffer(buf, dtype=np.uint8, count=32 * n_nodes, offset=off).copy(); off += 32 * n_nodes
    s.n_nodes = n_nodes
    s.tri_verts = np.frombuffer(buf, dtype=np.float32, count=9 * n_prims, offset=off).copy(); off += 36 * n_prims
    s.prim_material = np.frombuffer(buf, dtype=np.uint32, count=n_prims, offset=off).copy(); off += 4 * n_prims
    s.materials = (A.Material * n_mat).from_buffer_copy(buf, off); off += C.sizeof(A.Material) * n_mat
    s.lights = (A.Light * n_lights).from_buffer_copy(buf, off); off += C.sizeof(A.Light) * n_lights
    if has_medium:
        m = A.Medium()
        m.type = struct.unpack_from("<i", buf, off)[0]; off += 4
        vals = np.frombuffer(buf, dtype=np.float32, count=16 + 3 + 3 + 30 * 3 + 1, offset=off); off += 4 * len(vals)
        for i in range(16):
            m.world_to_volume[i] = float(vals[i])
        for i in range(3):
            m.p0[i] = float(vals[16 + i]); m.p1[i] = float(vals[19 + i])
        for i in range(A.NSPEC):
            m.sigma_a[i] = float(vals[22 + i]); m.sigma_s[i] = float(vals[52 + i]); m.le[i] = float(vals[82 + i])
        m.g = float(vals[112])
        m.nx, m.ny, m.nz = struct.unpack_from("<3i", buf, off); off += 12
        if m.type in (A.MEDIUM_GRID, A.MEDIUM_EXPONENTIAL):          # exponential: {a, b, updir.xyz} in the density slot
            cnt = m.nx * m.ny * m.nz
            s.density = np.frombuffer(buf, dtype=np.float32, count=cnt, offset=off).copy(); off += 4 * cnt
        s.medium = m
    if n_spheres:
        s.prim_shape = np.frombuffer(buf, dtype=np.uint32, count=n_prims, offset=off).copy(); off += 4 * n_prims
        s.spheres = (A.Sphere * n_spheres).from_buffer_copy(buf, off); off += C.sizeof(A.Sphere) * n_spheres
    if n_light_tris:
        s.light_tris = np.frombuffer(buf, dtype=np.float32, count=9 * n_light_tris, offset=off).reshape(-1, 9).copy(); off += 36 * n_light_tris
    return s


def write_scene(path, s):
    with open(path, "wb") as f:
        f.write(b"PVSCN001")
        f.write(struct.pack("<8I", s.n_nodes, s.n_prims, len(s.materials), len(s.lights), 1 if s.medium is not None else 0, len(s.spheres),
                            0 if s.light_tris is None else len(np.asarray(s.light_tris).reshape(-1, 9)), 0))
        f.write(np.asarray(s.world_bound, dtype=np.float32).tobytes())
        f.write(np.asarray(s.cie_y, dtype=np.float32).tobytes())
        f.write(np.asarray(s.nodes, dtype=np.uint8).tobytes())
        f.write(np.asarray(s.tri_verts, dtype=np.float32).tobytes())
        f.write(np.asarray(s.prim_material, dtype=np.uint32).tobytes())
        f.write(bytes(s.materials)); f.write(bytes(s.lights))
        if s.medium is not None:
            m = s.medium
            f.write(struct.pack("<i", m.type))
            f.write(np.array(list(m.world_to_volume) + list(m.p0) + list(m.p1) + list(m.sigma_a) + list(m.sigma_s)
                             + list(m.le) + [m.g], dtype=np.float32).tobytes())
            f.write(struct.pack("<3i", m.nx, m.ny, m.nz))
            if m.type in (A.MEDIUM_GRID, A.MEDIUM_EXPONENTIAL):
                f.write(np.asarray(s.density, dtype=np.float32).tobytes())
        if len(s.spheres):
            f.write(np.asarray(s.prim_shape, dtype=np.uint32).tobytes()); f.write(bytes(s.spheres))
        if s.light_tris is not None and len(s.light_tris):
            f.write(np.asarray(s.light_tris, dtype=np.float32).tobytes())


def _hdr(magic, n):
    return magic + struct.pack("<Q", n)


def write_photons(path, pos, wi, alpha):
    n = len(pos)
    rec = np.concatenate([np.asarray(pos, np.float32).reshape(n, 3), np.asarray(wi, np.float32).reshape(n, 3),
                          np.asarray(alpha, np.float32).reshape(n, A.NSPEC)], axis=1)
    with open(path, "wb") as f:
        f.write(_hdr(b"PVPHOT01", n)); f.write(np.ascontiguousarray(rec).tobytes())


def read_photons(path):
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"PVPHOT01"
    n = struct.unpack_from("<Q", buf, 8)[0]
    rec = np.frombuffer(buf, dtype=np.float32, count=36 * n, offset=16).reshape(n, 36)
    return (np.ascontiguousarray(rec[:, 0:3]), np.ascontiguousarray(rec[:, 3:6]), np.ascontiguousarray(rec[:, 6:36]))


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("mint", np.float32), ("maxt", np.float32),
                      ("time", np.float32), ("u_scatter", np.float32)])


def make_rays(o, d, mint=0.0, maxt=np.inf, time=0.0, u_scatter=0.5):
    o = np.asarray(o, np.float32).reshape(-1, 3)
    r = np.zeros(len(o), dtype=RAY_DTYPE)
    r["o"] = o; r["d"] = np.asarray(d, np.float32).reshape(-1, 3)
    r["mint"] = mint; r["maxt"] = maxt; r["time"] = time; r["u_scatter"] = u_scatter
    return r


def write_rays(path, rays):
    with open(path, "wb") as f:
        f.write(_hdr(b"PVRAY001", len(rays))); f.write(np.ascontiguousarray(rays).tobytes())


def write_queries(path, pts, w):
    n = len(pts)
    rec = np.concatenate([np.asarray(pts, np.float32).reshape(n, 3), np.asarray(w, np.float32).reshape(n, 3)], axis=1)
    with open(path, "wb") as f:
        f.write(_hdr(b"PVQRY001", n)); f.write(np.ascontiguousarray(rec).tobytes())


def read_knn(path):
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"PVKNN001"
    n = struct.unpack_from("<Q", buf, 8)[0]
    k = struct.unpack_from("<I", buf, 16)[0]
    rec = np.frombuffer(buf, dtype=np.uint32, count=n * (1 + 2 * k), offset=20).reshape(n, 1 + 2 * k)
    return rec[:, 0].copy(), rec[:, 1:1 + k].copy(), rec[:, 1 + k:].copy().view(np.float32)


def read_spectra(path, magic, per=1):
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == magic, (buf[:8], magic)
    n = struct.unpack_from("<Q", buf, 8)[0]
    return np.frombuffer(buf, dtype=np.float32, count=n * per * A.NSPEC, offset=16).reshape(n, per, A.NSPEC).copy()


def read_hits(path):
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"PVHIT001"
    n = struct.unpack_from("<Q", buf, 8)[0]
    rec = np.frombuffer(buf, dtype=np.uint32, count=3 * n, offset=16).reshape(n, 3)
    return rec[:, 0].copy(), rec[:, 1].copy().view(np.float32), rec[:, 2].copy()
