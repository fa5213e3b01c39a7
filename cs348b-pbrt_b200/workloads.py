"""Synthetic workloads of BASELINE.json (configs 2, 3, 5) as data: scene + photon set + camera rays.

Spectra (sigma_a, sigma_s, light I, Kd -> 30 bins) come from the reference's own RGB->spectrum
conversion through the scene fixtures exported by oracle/ref_harness (tests/golden/*.scn); the
density grid, photons and rays are closed form / seeded, generated with numpy on the host so both
bench arms (CUDA path and CPU reference) see identical inputs.
"""
import os
import numpy as np
from . import sceneio, scenes

_HERE = os.path.dirname(os.path.abspath(__file__))
_GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")

CONFIGS = {
    # name: grid n, photons, xres, yres, stepsize, nused, maxdist
    "config3": dict(grid=256, photons=16_000_000, xres=1920, yres=1080, stepsize=2.0 / 64, nused=512, maxdist=0.018,
                    label="heterogeneous 256^3 density-grid medium, 16M photons, 1920x1080, fixed-radius gather"),
    "config5": dict(grid=256, photons=128_000_000, xres=3840, yres=2160, stepsize=2.0 / 64, nused=512, maxdist=0.009,
                    label="stress: 128M photons, 3840x2160, 64 ray-march steps/ray"),
    "config2": dict(grid=0, photons=1_000_000, xres=512, yres=512, stepsize=0.05, nused=50, maxdist=0.25,
                    label="synthetic Cornell box + homogeneous medium, 1M volume photons, 512x512, k=50 gather"),
    "tiny": dict(grid=32, photons=200_000, xres=160, yres=90, stepsize=2.0 / 64, nused=512, maxdist=0.08,
                 label="reduced config3 (tests)"),
}


def load_scene(cfg):
    if cfg["grid"]:
        s = sceneio.read_scene(os.path.join(_GOLDEN, "cornell_grid32.scn"))
        n = cfg["grid"]
        s.density = scenes.blob_density(n)
        s.medium.nx = s.medium.ny = s.medium.nz = n
        return s
    return sceneio.read_scene(os.path.join(_GOLDEN, "cornell_homog.scn"))


PHOTON_CHUNK = 1_000_000


def photons_from_density(scene, n, seed=0x5EED, lo=0, hi=None):
    """Photons lo..hi of an n-photon set whose positions are distributed like the medium density (piecewise
    constant per voxel: inverse CDF over voxels + jitter), wi uniform on the sphere, alpha a near-flat spectrum of
    weight ~1/n.  Generated in chunks of PHOTON_CHUNK, each from its own seeded stream, so any rank can produce any
    slice (lo and hi must be chunk-aligned or the end of the set) and the union does not depend on the sharding."""
    hi = n if hi is None else hi
    assert lo % PHOTON_CHUNK == 0 and (hi % PHOTON_CHUNK == 0 or hi == n)
    m = hi - lo
    pos = np.empty((m, 3), np.float32); wi = np.empty((m, 3), np.float32); alpha = np.empty((m, 30), np.float32)
    cdf = None
    if scene.density is not None:
        g = scene.medium.nx
        cdf = np.cumsum(scene.density.astype(np.float64))
        cdf /= cdf[-1]
    base = (1.0 + 0.01 * np.arange(30, dtype=np.float32))[None, :] / np.float32(n)
    for a in range(lo, hi, PHOTON_CHUNK):
        b = min(hi, a + PHOTON_CHUNK)
        rng = np.random.default_rng([seed, a // PHOTON_CHUNK])
        k = b - a
        if cdf is not None:
            v = np.searchsorted(cdf, rng.random(k), side="right").astype(np.int64)
            np.minimum(v, g * g * g - 1, out=v)
            z, rem = np.divmod(v, g * g)
            y, x = np.divmod(rem, g)
            j = rng.random((k, 3))
            p = np.stack([(x + j[:, 0]) / g * 2 - 1, (y + j[:, 1]) / g * 2 - 1, (z + j[:, 2]) / g * 2 - 1], axis=1)
        else:
            p = rng.uniform(-1, 1, size=(k, 3))
        pos[a - lo:b - lo] = np.clip(p, -0.999999, 0.999999).astype(np.float32)
        zc = rng.uniform(-1, 1, size=k); phi = rng.uniform(0, 2 * np.pi, size=k)
        r = np.sqrt(np.maximum(0, 1 - zc * zc))
        wi[a - lo:b - lo] = np.stack([r * np.cos(phi), r * np.sin(phi), zc], axis=1).astype(np.float32)
        alpha[a - lo:b - lo] = base * rng.uniform(0.5, 1.5, size=(k, 1)).astype(np.float32)
    return pos, wi, alpha


def photon_slice(n, rank, world):
    """Chunk-aligned slice [lo, hi) of an n-photon set for rank `rank` of `world`."""
    nchunks = (n + PHOTON_CHUNK - 1) // PHOTON_CHUNK
    c0 = nchunks * rank // world; c1 = nchunks * (rank + 1) // world
    return min(n, c0 * PHOTON_CHUNK), min(n, c1 * PHOTON_CHUNK)


def tile_order(xres, yres, tile=8):
    """Pixel order in tile x tile blocks (row-major tiles), like the reference's image tiles
    (renderers/samplerrenderer.cpp:206-217, core/sampler.cpp:55-74): neighbouring rays stay neighbours."""
    ys, xs = np.meshgrid(np.arange(yres), np.arange(xres), indexing="ij")
    key = ((ys // tile) * ((xres + tile - 1) // tile) + xs // tile) * (tile * tile) + (ys % tile) * tile + xs % tile
    return np.argsort(key.reshape(-1), kind="stable")


def deal_tiles(rays, order, world, tile=8, group=8, medium_box=((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)), xres=None, density=None):
    """Which rank gathers which rays (image tiles of camera rays are sharded).  The unit dealt is a block of
    group x group tiles (64 x 64 pixels): big enough that a rank's march steps stay spatially dense -- the cell-batched gather
    shares one staged block of photon cells among 32 neighbouring steps, and single 8 x 8 tiles dealt round-robin thin the steps
    of a rank out (measured at 2 ranks: 340 instead of 291 distance tests per lookup) -- and dealt by COST, not round-robin.  The
    cost of a ray is what its march steps will cost: its length inside the medium's bounding box (the number of steps, up to the
    constant step size) times (0.4 + the medium's density along it relative to the mean density) -- a lookup scans and sums photons
    in proportion to the local photon density, which follows the medium's -- plus a fixed part per ray.  Blocks go to the
    least-loaded rank in order of decreasing cost (LPT).  Same answer on every rank.  Returns rank_of_ray, aligned with `order`."""
    o = rays["o"][order].astype(np.float32); d = rays["d"][order].astype(np.float32)
    lo, hi = np.asarray(medium_box[0], np.float32), np.asarray(medium_box[1], np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (lo - o) / d; t1 = (hi - o) / d
    tn = np.maximum(np.nanmax(np.minimum(t0, t1), axis=1), 0.0); tf = np.nanmin(np.maximum(t0, t1), axis=1)
    chord = np.clip(tf - tn, 0.0, None)
    weight = np.ones_like(chord)
    if density is not None:
        g = int(round(len(density) ** (1.0 / 3.0)))
        dens = np.asarray(density, np.float32).reshape(g, g, g)           # [z][y][x]
        acc = np.zeros(len(chord), np.float32)
        K = 12
        for k in range(K):                                                 # nearest-voxel density at K points of the chord
            t = tn + (k + 0.5) / K * chord
            p = o + d * t[:, None]
            idx = np.clip(((p - lo) / (hi - lo) * g).astype(np.int32), 0, g - 1)
            acc += dens[idx[:, 2], idx[:, 1], idx[:, 0]]
        weight = 0.4 + acc / K / max(float(dens.mean()), 1e-30)
    pix = order
    gx = (pix % xres) // (tile * group); gy = (pix // xres) // (tile * group)
    ngx = (xres + tile * group - 1) // (tile * group)
    block = gy * ngx + gx
    nblocks = int(block.max()) + 1
    cost = np.bincount(block, weights=chord * weight + 0.1, minlength=nblocks)       # + a fixed part per ray (march set-up, the 272 B of L and T)
    load = np.zeros(world)
    owner = np.zeros(nblocks, np.int64)
    for b in np.argsort(-cost, kind="stable"):
        r = int(np.argmin(load)); owner[b] = r; load[r] += cost[b] + 1e-9
    return owner[block]


def frame_rays(cfg, rank=0, world=1, tile=8, density=None):
    """Camera rays of the frame in tile order; with several ranks, blocks of 8 x 8 tiles are dealt by cost (deal_tiles; `density`
    = the medium's density grid, if it has one).  Returns (rays, global ray indices)."""
    rays = scenes.camera_rays(cfg["xres"], cfg["yres"])
    order = tile_order(cfg["xres"], cfg["yres"], tile)
    if world > 1:
        owner = deal_tiles(rays, order, world, tile=tile, xres=cfg["xres"], density=density)
        order = order[owner == rank]
    return np.ascontiguousarray(rays[order]), order


def gather_bytes(stats, nrays, b_ph=144):
    """Algorithmic bytes of one gather pass: sum_lookups nFound*B_ph + lookups*28 + rays*272."""
    return stats.photons_found * b_ph + stats.lookups * 28 + nrays * 272
