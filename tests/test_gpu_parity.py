"""GPU (-m gpu): the CUDA path, called through the C ABI (csrc/libpv.so via ctypes), against
 (a) golden vectors produced by the real reference (tests/golden/*.npz) and
 (b) the pinned CPU oracle on the same seeded inputs.
Bars: k-NN index sets, distances, BVH hit ids and hit distances BIT-EXACT; radiance within 1e-4 relative
(north_star), stated per test."""
import numpy as np
import pytest
import oracle_lib as O

pytestmark = pytest.mark.gpu
RTOL_RADIANCE = 1e-4
# rainbow_vol / prism_small = BASELINE configs 1 and 4 (the reference project's own scenes, reduced counts; tests/golden/make_golden.py)
# sphere_glass / sphere_disp: Sphere primitives (the project's glass-ball scene, and a rotated / scaled / partial / dispersive variant)
ALL_SCENES = ["cornell_homog", "cornell_grid32", "rainbow_vol", "prism_small", "sphere_glass", "sphere_disp", "cornell_exp"]


def relerr(a, b, floor=1e-30):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)




def test_library_is_the_cuda_one(pkg):
    import os
    assert os.path.exists(pkg.library_path())
    pv = pkg.PhotonVolume(device=0)
    assert pv.lib.pv_version() >= 100
    pv.close()


@pytest.mark.parametrize("name", ALL_SCENES)
def test_knn_bit_exact_vs_reference(golden, pv_factory, name):
    g, scene = golden(name)
    pv = pv_factory(nused=int(g["params"][0]), maxdist=float(g["params"][1]))
    pv.set_scene(scene)
    pv.set_photons(g["shot_pos"], g["shot_wi"], g["shot_alpha"])
    pv.build()
    for k in (50, 16, 100):
        if "knn%d_idx" % k not in g:
            continue
        nf, idx, d2 = pv.Lookup(g["q_pts"], k=k, r2=float(g["knn%d_r2" % k][0]))
        assert np.array_equal(nf, g["knn%d_nfound" % k])
        assert np.array_equal(idx, g["knn%d_idx" % k])
        assert np.array_equal(d2.view(np.uint32), g["knn%d_d2" % k].view(np.uint32))


@pytest.mark.parametrize("build_hint", [(0.25, 50), (0.05, 8), (1.0, 1), (0.5, 300)])
def test_knn_synthetic_bit_exact_any_cell_size(golden, pv_factory, build_hint):
    """The result must not depend on the grid the map was built for."""
    g, _ = golden("synthetic_knn")
    pv = pv_factory()
    n = len(g["pos"])
    alpha = np.full((n, 30), 1.0 / n, np.float32)
    pv.set_photons(g["pos"], g["wi"], alpha)
    pv.build(maxdist=build_hint[0], nused=build_hint[1])
    for k in (50, 8, 300, 64, 1):
        nf, idx, d2 = pv.Lookup(g["q_pts"], k=k, r2=float(g["knn%d_r2" % k][0]))
        assert np.array_equal(nf, g["knn%d_nfound" % k]), k
        assert np.array_equal(idx, g["knn%d_idx" % k]), k
        assert np.array_equal(d2.view(np.uint32), g["knn%d_d2" % k].view(np.uint32)), k


def test_knn_ties_broken_by_photon_index(pv_factory):
    """Duplicate positions => exact distance ties straddling the k-th place: lower photon index wins."""
    rng = np.random.default_rng(3)
    base = rng.uniform(-1, 1, size=(400, 3)).astype(np.float32)
    pos = np.concatenate([base, base, base])            # every position three times
    n = len(pos)
    pv = pv_factory()
    pv.set_photons(pos, np.tile([[0, 0, 1]], (n, 1)).astype(np.float32), np.ones((n, 30), np.float32))
    pv.build(maxdist=0.5, nused=10)
    pts = base[:64] + np.float32(1e-3)
    for k in (1, 2, 4, 10, 31):
        nf, idx, d2 = pv.Lookup(pts, k=k, r2=0.25)
        bnf, bidx, bd2 = O.knn_brute(pos, pts, k, 0.25)
        assert np.array_equal(nf, bnf) and np.array_equal(idx, bidx), k
        assert np.array_equal(d2.view(np.uint32), bd2.view(np.uint32)), k


def test_knn_edge_cases(pv_factory):
    pv = pv_factory()
    # empty map
    pv.set_photons(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 30), np.float32))
    pv.build(maxdist=0.1, nused=5)
    nf, idx, d2 = pv.Lookup(np.zeros((3, 3), np.float32), k=5, r2=0.01)
    assert (nf == 0).all() and (idx == 0xFFFFFFFF).all() and np.isinf(d2).all()
    # fewer photons than k, queries far outside the grid, a single photon
    pos = np.array([[0.25, 0.5, 0.75]], np.float32)
    pv.set_photons(pos, np.array([[0, 0, 1]], np.float32), np.ones((1, 30), np.float32))
    pv.build(maxdist=0.125, nused=5)
    pts = np.array([[0.25, 0.5, 0.75], [0.3125, 0.5, 0.75], [5, 5, 5], [0.25, 0.5, 0.875]], np.float32)
    nf, idx, d2 = pv.Lookup(pts, k=5, r2=0.015625)
    bnf, bidx, bd2 = O.knn_brute(pos, pts, 5, 0.015625)
    assert np.array_equal(nf, bnf) and np.array_equal(idx, bidx)
    assert nf.tolist() == [1, 1, 0, 0]          # d2 == r2 (0.125^2 exactly) is excluded: strict <, kdtree.h:180
    # zero queries
    nf, idx, d2 = pv.Lookup(np.zeros((0, 3), np.float32), k=5, r2=0.01)
    assert len(nf) == 0


def test_lookup_before_build_fails_loudly(pkg, pv_factory):
    pv = pv_factory()
    pv.set_photons(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32), np.zeros((4, 30), np.float32))
    with pytest.raises(pkg.PVError) as e:
        pv.Lookup(np.zeros((1, 3), np.float32), k=1, r2=1.0)
    assert "not built" in str(e.value)


@pytest.mark.parametrize("name", ALL_SCENES)
def test_bvh_hits_bit_exact_vs_reference(golden, pv_factory, name):
    g, scene = golden(name)
    pv = pv_factory()
    pv.set_scene(scene)
    prim, t = pv.Intersect(g["hit_rays"])
    assert np.array_equal(prim, g["hit_prim"])
    assert np.array_equal(t.view(np.uint32), g["hit_t"].view(np.uint32))
    occ = pv.IntersectP(g["hit_rays"])
    assert np.array_equal(occ.astype(np.uint32), g["hit_occluded"])


@pytest.mark.parametrize("name", ALL_SCENES)
def test_transmittance_vs_reference(golden, pv_factory, name):
    g, scene = golden(name)
    pv = pv_factory(stepsize=float(g["params"][2]))
    pv.set_scene(scene)
    T = pv.Transmittance(g["li_rays"], g["tr_u"])          # step = 4*stepsize, the sample == NULL branch
    assert relerr(T, g["tr_T"]).max() < 1e-5                # libm expf differs in the last ulps


@pytest.mark.parametrize("name", ALL_SCENES)
def test_lphoton_vs_reference(golden, pv_factory, name):
    g, scene = golden(name)
    pv = pv_factory(nused=int(g["params"][0]), maxdist=float(g["params"][1]))
    pv.set_scene(scene)
    pv.set_photons(g["shot_pos"], g["shot_wi"], g["shot_alpha"])
    pv.build()
    L = pv.LPhoton(g["q_pts"], g["q_w"])
    ref = g["lphoton_L"]
    assert np.array_equal(L == 0, ref == 0)                 # the nFound < 10 rule hits the same queries
    assert relerr(L, ref)[ref > 0].max() < RTOL_RADIANCE


@pytest.mark.parametrize("name", ["cornell_homog", "rainbow_vol"])
def test_li_homogeneous_vs_reference(golden, pv_factory, name):
    """Homogeneous (or rainbow) medium + one delta light: Li does not depend on any random draw, so the CUDA result is
    compared directly with what the reference binary returned."""
    g, scene = golden(name)
    pv = pv_factory(stepsize=float(g["params"][2]), nused=int(g["params"][0]), maxdist=float(g["params"][1]))
    pv.set_scene(scene)
    pv.set_photons(g["shot_pos"], g["shot_wi"], g["shot_alpha"])
    pv.build()
    L, T = pv.Li(g["li_rays"])
    assert relerr(T, g["li_T"]).max() < 1e-5
    m = g["li_L"] > 0
    assert m.any()
    assert relerr(L, g["li_L"])[m].max() < RTOL_RADIANCE
    assert np.array_equal(L == 0, g["li_L"] == 0)


@pytest.mark.parametrize("name,flags", [("cornell_homog", 0), ("cornell_grid32", 0), ("cornell_grid32", 1), ("cornell_grid32", 2),
                                        ("rainbow_vol", 0), ("prism_small", 0), ("sphere_glass", 0), ("sphere_disp", 0), ("cornell_exp", 0)])
def test_li_vs_oracle_same_philox_stream(golden, pv_factory, name, flags):
    g, scene = golden(name)
    stepsize, nused, maxdist = float(g["params"][2]), int(g["params"][0]), float(g["params"][1])
    pv = pv_factory(stepsize=stepsize, nused=nused, maxdist=maxdist, seed=0xC0FFEE)
    pv.set_scene(scene)
    pv.set_photons(g["shot_pos"], g["shot_wi"], g["shot_alpha"])
    pv.build()
    rays = g["li_rays"]
    L, T = pv.Li(rays, ray_index_base=1000, flags=flags)
    tree = O.KdTree(g["shot_pos"])
    oL, oT, ost = O.gather(scene, tree, g["shot_wi"], g["shot_alpha"], rays, stepsize, nused, maxdist, seed=0xC0FFEE,
                           ray_index_base=1000, flags=flags)
    assert relerr(T, oT).max() < 1e-5
    m = oL > 0
    assert m.any()
    assert relerr(L, oL)[m].max() < RTOL_RADIANCE
    st = pv.gather_stats(reset=True)
    if not (flags & 2):
        assert st.lookups == ost.lookups and st.photons_found == ost.photons_found


def test_li_rays_missing_the_medium(golden, pv_factory):
    g, scene = golden("cornell_homog")
    pv = pv_factory(stepsize=0.05, nused=50, maxdist=0.25)
    pv.set_scene(scene)
    pv.set_photons(g["shot_pos"], g["shot_wi"], g["shot_alpha"])
    pv.build()
    rays = pkg_rays(np.array([[0, 5, -5], [0, 0, -3]], np.float32), np.array([[0, 0, 1], [0, 0, 1]], np.float32))
    rays["maxt"][1] = 1.5                       # stops before the box
    L, T = pv.Li(rays)
    assert (L == 0).all() and (T == 1).all()
    L, T = pv.Li(rays[:0])
    assert L.shape == (0, 30)


def pkg_rays(o, d):
    from __graft_entry__ import load_package
    return load_package().sceneio.make_rays(o, d)


@pytest.mark.parametrize("name,wanted,sstep", [("cornell_homog", 3000, 0.05), ("cornell_grid32", 1200, 0.05), ("rainbow_vol", 1500, 0.1),
                                               ("prism_small", 4000, 0.1), ("sphere_glass", 4000, 0.1), ("sphere_disp", 4000, 0.1),
                                               ("cornell_exp", 1000, 0.05)])
def test_shooter_vs_oracle_same_philox_stream(golden, pv_factory, name, wanted, sstep):
    """Same per-path Philox streams on both sides: photons are matched one to one by (path, deposit ordinal).
    prism_small: every path goes through the dispersive glass wedge (splitSpectrum into 30 monochromatic photons, Cauchy refraction)."""
    g, scene = golden(name)
    istep = float(g["params"][2])
    pv = pv_factory(stepsize=istep, seed=77)
    pv.set_scene(scene)
    st = pv.Preprocess(wanted, stepsize=sstep, max_photon_depth=5, build=False)
    pos, wi, alpha, ids = pv.get_photons()
    ref = O.shoot(scene, wanted, sstep, istep, seed=77, rng_mode=O.PHILOX, nthreads=8)
    assert ref["rc"] == 0
    assert st.paths == ref["nshot"]
    assert st.stack_overflows == 0
    # one to one on ids.  Measured (tools/shoot_mismatch.py, all eight scenes): the id lists are IDENTICAL, positions agree to 3e-6
    # (3e-5 on the spheres), weights to 1e-6 -- libm vs CUDA differences (sincosf, expf, acosf) are there but flip no decision on
    # these streams.  The bars are those measurements with a margin, not a statistical allowance.
    assert np.array_equal(ids, ref["ids"])
    dpos = np.abs(pos - ref["pos"]).max(axis=1)
    assert dpos.max() < 1e-4 and np.quantile(dpos, 0.99) < 5e-6
    e = relerr(alpha, ref["alpha"]).max(axis=1)
    if name.startswith("sphere"):
        # curved glass: the normal comes out of acosf / sinf / atan2f (libm vs CUDA, last-ulp differences), and the Fresnel term
        # has a square-root singularity at the critical angle that internal reflections in the ball do reach -- a 1e-7 change
        # of the normal moves F by ~sqrt(1e-7) there.  99 % of the photons agree to 1e-5, the worst to 1.5 per cent.
        assert np.quantile(e, 0.99) < 1e-5 and e.max() < 5e-2, (np.quantile(e, 0.99), e.max())
    else:
        assert e.max() < 1e-5
    assert np.all(np.diff(ids.astype(np.int64)) > 0)        # deterministic order: sorted by (path, ordinal)


def test_shooter_lambda_is_path_state_in_the_underflow_regime(golden, pv_factory):
    """The regime that separates `Spectrum::lambda as path state` from `lambda inferred from the bins` (tests/golden/underflow_glass.*,
    see test_oracle_golden.py): dispersive wedge in a dense exponential medium, photon weights underflow to zero between glass
    faces.  GPU vs the pinned oracle on the same per-path Philox streams: identical id lists -- black photons included."""
    g, scene = golden("underflow_glass")
    wanted, sstep, istep = 3000, float(g["params"][1]), float(g["params"][2])
    pv = pv_factory(stepsize=istep, seed=113)
    pv.set_scene(scene)
    st = pv.Preprocess(wanted, stepsize=sstep, max_photon_depth=5, build=False)
    pos, wi, alpha, ids = pv.get_photons()
    ref = O.shoot(scene, wanted, sstep, istep, seed=113, rng_mode=O.PHILOX, nthreads=8)
    assert ref["rc"] == 0 and st.paths == ref["nshot"] and st.stack_overflows == 0
    black_ref = ref["alpha"].max(axis=1) == 0
    assert black_ref.sum() >= 5                                  # the regime is reached
    assert np.array_equal(ids, ref["ids"])                       # identical id lists (measured; a re-split or cut-short path would change them)
    assert np.abs(pos - ref["pos"]).max() < 1e-5
    assert relerr(alpha, ref["alpha"], floor=1e-38).max() < 1e-4
    # the same photons are black on both sides
    bg = alpha.max(axis=1) == 0
    assert np.array_equal(bg, black_ref) and bg.sum() >= 5


def test_shooter_sharded_blocks_reproduce_single_rank(golden, pv_factory, pkg):
    """Emission sharded by block over 2 'ranks' (run back to back on one GPU) gives the same photon set."""
    import ctypes as C
    g, scene = golden("cornell_homog")
    A = pkg._abi
    pv1 = pv_factory(stepsize=0.05, seed=5)
    pv1.set_scene(scene)
    st = pv1.Preprocess(1500, stepsize=0.05, build=False)
    p1, w1, a1, i1 = pv1.get_photons()
    nblocks = int(st.blocks)
    parts = []
    for rank in range(2):
        pv = pv_factory(stepsize=0.05, seed=5)
        pv.set_scene(scene)
        prm = A.ShootParams(0.05, 0.05, 5, 5, rank, 2, 0, 0.0)
        counts = (C.c_uint32 * nblocks)()
        s2 = A.ShootStats()
        pv._chk(pv.lib.pv_shoot_blocks(pv.ctx, C.c_uint64(1), C.c_uint32(nblocks), C.byref(prm), counts, C.byref(s2)))
        pv._chk(pv.lib.pv_shoot_finish(pv.ctx, C.c_uint64(nblocks)))
        parts.append(pv.get_photons())
    ids = np.concatenate([p[3] for p in parts]); order = np.argsort(ids)
    assert np.array_equal(ids[order], i1)
    assert np.array_equal(np.concatenate([p[0] for p in parts])[order], p1)
    assert np.array_equal(np.concatenate([p[2] for p in parts])[order], a1)


# ---- surface photon maps
SURF_SCENES = ["cornell_surf", "cornell_surf_disp", "rainbow_surf"]
MAP_KEYS = ("volume", "caustic", "indirect", "direct", "radiance")


@pytest.mark.gpu
@pytest.mark.parametrize("name", SURF_SCENES)
def test_all_maps_vs_oracle_same_philox_stream(golden, pv_factory, pkg, name):
    """pv_shoot_maps against the pinned oracle on the same per-path Philox streams: every class matched one to one on
    (class, path, deposit ordinal), same nshot and per-map path counts (i.e. the done flags flipped at the same blocks)."""
    g, scene = golden(name)
    nv, nc, ni, fg, sstep, istep = g["params"][:6]
    pv = pv_factory(stepsize=float(istep), seed=91)
    pv.set_scene(scene)
    st = pv.PreprocessMaps(int(nv), int(nc), int(ni), bool(fg), stepsize=float(sstep), max_photon_depth=5)
    ref = O.shoot_maps(scene, int(nv), int(nc), int(ni), bool(fg), float(sstep), float(istep), seed=91, rng_mode=O.PHILOX)
    assert ref["rc"] == 0
    assert st.shoot.stack_overflows == 0
    assert (st.nshot, st.n_caustic_paths, st.n_indirect_paths, st.n_direct_paths) == (ref["nshot"], ref["caustic_paths"], ref["indirect_paths"],
                                                                                    ref["direct_paths"])
    assert abs(int(st.n_volume_paths) - int(ref["volume_paths"])) <= 0.005 * ref["volume_paths"] + 2
    for which, key in enumerate(MAP_KEYS):
        pos, wi, alpha, ids = pv.get_map_photons(which)
        r = ref[key]
        assert len(ids) == int(st.n[which])
        if len(r["ids"]) == 0:
            assert len(ids) == 0, key
            continue
        assert np.all((ids >> np.uint64(60)) == which), key
        assert np.all(np.diff(ids.astype(np.int64)) > 0), key              # ordered by (path, deposit ordinal)
        common, ia, ib = np.intersect1d(ids, r["ids"], return_indices=True)
        assert len(common) >= 0.995 * max(len(ids), len(r["ids"])), key
        assert abs(len(ids) - len(r["ids"])) <= 0.005 * len(r["ids"]) + 2, key
        dpos = np.abs(pos[ia] - r["pos"][ib]).max(axis=1)
        assert np.quantile(dpos, 0.99) < 1e-4, key
        ok = dpos < 1e-4
        assert np.abs(wi[ia][ok] - r["wi"][ib][ok]).max() < 1e-4, key
        assert relerr(alpha[ia][ok], r["alpha"][ib][ok]).max() < 1e-3, key


def ephoton_rule(maps, counts, sites, normals, rho_r, k, r2):
    """Brute-force EPhoton / ComputeRadianceTask with the rule the CUDA path states: the k smallest by (d2, photon index) among
    d2 < r2, d2 in the reference's unfused fp32 ((dx*dx + dy*dy) + dz*dz).  Also returns the sites where some map has a TIE at
    the k-th distance: there the reference's own result depends on its kd-tree traversal order (strict `<` in
    KdTree::Lookup, heap ties by node address), so it is not a parity target."""
    sites = np.asarray(sites, np.float32); n = len(sites)
    E = np.zeros((n, 30), np.float64); tie = np.zeros(n, bool)
    r2 = np.float32(r2)
    for (pos, wi, alpha), count in zip(maps, counts):
        if len(pos) == 0 or count == 0:
            continue
        pos = np.asarray(pos, np.float32)
        d = sites[:, None, :] - pos[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]          # float32, same association
        for i in range(n):
            cand = np.nonzero(d2[i] < r2)[0]
            if len(cand) == 0:
                continue
            order = cand[np.lexsort((cand, d2[i][cand]))]
            md2 = r2
            if len(order) >= k:
                if len(order) > k and d2[i][order[k]] == d2[i][order[k - 1]]:
                    tie[i] = True
                order = order[:k]
                md2 = d2[i][order].max()
            facing = (np.asarray(wi, np.float32)[order] * normals[i]).sum(axis=1) > 0
            den = np.float32(np.float64(np.float32(np.float32(count) * md2)) * np.pi)
            E[i] += np.asarray(alpha, np.float64)[order][facing].sum(axis=0) / np.float64(den)
    black = ~(np.asarray(rho_r) != 0).any(axis=1)
    Lo = (np.float32(1.0 / np.pi) * np.asarray(rho_r, np.float64)) * E
    Lo[black] = 0
    return Lo, tie


@pytest.mark.gpu
@pytest.mark.parametrize("name", SURF_SCENES)
def test_radiance_photons_vs_reference(golden, pv_factory, pkg, name):
    """pv_radiance_photons on the REFERENCE's own direct / indirect / caustic lists and radiance-photon sites (three grid builds +
    EPhoton lookups on the GPU): Lo within 1e-4 of the reference's ComputeRadianceTask at every site whose n_lookup-nearest sets are
    unique, and within 1e-4 of the brute-force rule (ties by photon index) everywhere.  Exact ties are real here: the 30
    monochromatic children of a photon that a dispersive glass face REFLECTS all land on the same point."""
    A = pkg._abi
    g, scene = golden(name)
    nlookup, md2 = int(g["params"][6]), float(g["params"][7])
    nshot, cp, ip, dp, vp = [int(x) for x in g["counts"]]
    pv = pv_factory()
    pv.set_scene(scene)
    for which, key in ((A.MAP_CAUSTIC, "caustic"), (A.MAP_INDIRECT, "indirect"), (A.MAP_DIRECT, "direct")):
        pv.set_map_photons(which, g[key + "_pos"], g[key + "_wi"], g[key + "_alpha"])
    pv.set_map_photons(A.MAP_RADIANCE, g["rad_pos"], g["rad_n"], g["rad_rho_r"])
    maps = [(g[k + "_pos"], g[k + "_wi"], g[k + "_alpha"]) for k in ("direct", "indirect", "caustic")]
    ref = g["rad_Lo"]
    assert (ref > 0).any()
    # (k = 8 would put md2 == 0 at sites where >= 8 reflected monochromatic photons coincide: the reference divides by zero there)
    for k, r2 in ((nlookup, md2), (40, md2), (20, 0.25 * md2)):
        got = pv.RadiancePhotons(k, r2, path_counts=(dp, ip, cp))
        rule, tie = ephoton_rule(maps, (dp, ip, cp), g["rad_pos"], g["rad_n"], g["rad_rho_r"], k, r2)
        # the reference itself for its own parameters, the pinned oracle (the reference's kd-tree order) for the others
        want = ref if k == nlookup else O.radiance(maps, [dp, ip, cp], g["rad_pos"], g["rad_n"], g["rad_rho_r"], k, r2)
        assert tie.mean() < 0.2, (k, tie.mean())
        u = ~tie
        assert relerr(got[u], want[u])[want[u] > 0].max() < 1e-4, k
        assert np.array_equal(got[u] == 0, want[u] == 0), k
        assert relerr(got, rule)[rule > 0].max() < 1e-4, k
        assert np.array_equal(got == 0, rule == 0), k


@pytest.mark.gpu
def test_all_maps_with_surface_maps_off_is_the_volume_pass(golden, pv_factory):
    g, scene = golden("cornell_homog")
    pv = pv_factory(stepsize=0.05, seed=5); pv.set_scene(scene)
    st = pv.Preprocess(1500, stepsize=0.05, build=False)
    a = pv.get_photons()
    pv2 = pv_factory(stepsize=0.05, seed=5); pv2.set_scene(scene)
    st2 = pv2.PreprocessMaps(1500, 0, 0, True, stepsize=0.05)
    b = pv2.get_photons()
    assert st2.nshot == st.paths
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert list(st2.n)[1:] == [0, 0, 0, 0]


@pytest.mark.gpu
def test_radiance_then_volume_gather_still_works(golden, pv_factory):
    """pv_radiance_photons rebuilds the grid per surface map; a pv_build afterwards restores the volume map."""
    g, scene = golden("cornell_surf")
    pv = pv_factory(stepsize=0.05, nused=50, maxdist=0.25, seed=3); pv.set_scene(scene)
    pv.PreprocessMaps(1500, 800, 2000, True, stepsize=0.05)
    pv.build()
    pts = pv.get_photons()[0][:64]
    nf0, idx0, d0 = pv.Lookup(pts)
    Lo = pv.RadiancePhotons(50, 0.0625)
    assert np.isfinite(Lo).all() and (Lo > 0).any()
    with pytest.raises(Exception):
        pv.Lookup(pts)                                                     # map not built any more: loud, not stale
    pv.build()
    nf1, idx1, d1 = pv.Lookup(pts)
    assert np.array_equal(nf0, nf1) and np.array_equal(idx0, idx1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell_surf", "rainbow_surf"])
def test_all_maps_sharded_over_two_ranks_reproduce_single_rank(golden, pv_factory, pkg, name):
    """pv_shoot_maps_ranks: emission dealt to 2 ranks by 4096-path blocks (two contexts on one GPU, driven by two threads whose
    all-reduce callback meets at a barrier).  Both ranks must take the same decisions (flag flips, roll-backs, last block) and the
    union of their photons, ordered by id, must be the single-rank result bit for bit -- for every photon class."""
    import threading
    from cs348b_pbrt_b200 import multigpu as MG
    g, scene = golden(name)
    nv, nc, ni, fg, sstep, istep = g["params"][:6]
    args = (int(nv), int(nc), int(ni), bool(fg))
    pv1 = pv_factory(stepsize=float(istep), seed=17); pv1.set_scene(scene)
    st1 = pv1.PreprocessMaps(*args, stepsize=float(sstep), max_photon_depth=5)
    world = 2
    pvs = []
    for r in range(world):
        p = pv_factory(stepsize=float(istep), seed=17); p.set_scene(scene); pvs.append(p)
    barrier = threading.Barrier(world)
    slots = [None] * world
    stats = [None] * world; errs = []

    def run(r):
        def allreduce(arr):
            slots[r] = arr.copy()
            barrier.wait(timeout=120)
            total = sum(s.astype(np.uint64) for s in slots).astype(np.uint32)
            barrier.wait(timeout=120)
            arr[:] = total
        try:
            stats[r] = pvs[r].PreprocessMapsRanks(*args, rank=r, world=world, allreduce=allreduce, stepsize=float(sstep), max_photon_depth=5)
        except Exception as e:          # pragma: no cover
            errs.append(e); barrier.abort()
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th: t.start()
    for t in th: t.join(timeout=300)
    assert not errs, errs
    for r in range(world):
        s = stats[r]
        assert (s.nshot, s.blocks, s.n_caustic_paths, s.n_indirect_paths, s.n_direct_paths, s.n_volume_paths) == (
            st1.nshot, st1.blocks, st1.n_caustic_paths, st1.n_indirect_paths, st1.n_direct_paths, st1.n_volume_paths)
    for which in range(5):
        want = pv1.get_map_photons(which)
        got = MG.merge_by_id([p.get_map_photons(which) for p in pvs])
        assert sum(int(stats[r].n[which]) for r in range(world)) == int(st1.n[which]) == len(want[3])
        for a, b in zip(got, want):
            assert np.array_equal(a, b), which
    assert min(int(stats[r].n[0]) for r in range(world)) > 0          # both ranks really traced photons


@pytest.mark.gpu
@pytest.mark.parametrize("name", SURF_SCENES)
def test_surface_integrator_lookups_vs_reference(golden, pv_factory, pkg, name):
    """pv_select_map + pv_surface_lphoton / pv_radiance_nearest against the reference's own LPhoton (integrators/photonmap.cpp:62-108)
    and RadiancePhotonProcess lookup (:238-243) on the reference's photon lists: Lr, Lt within 1e-4; nearest distance bit-exact, index
    equal wherever the nearest facing radiance photon is unique."""
    A = pkg._abi
    g, scene = golden(name)
    nlookup, md2 = int(g["params"][6]), float(g["params"][7])
    nshot, cp, ip, dp, vp = [int(x) for x in g["counts"]]
    pv = pv_factory(); pv.set_scene(scene)
    for which, key in ((A.MAP_CAUSTIC, "caustic"), (A.MAP_INDIRECT, "indirect"), (A.MAP_DIRECT, "direct")):
        pv.set_map_photons(which, g[key + "_pos"], g[key + "_wi"], g[key + "_alpha"])
    pv.set_map_photons(A.MAP_RADIANCE, g["rad_pos"], g["rad_n"], g["rad_rho_r"])
    inv_pi = np.float32(1.0 / np.pi)
    for which, key, npaths in ((A.MAP_CAUSTIC, "caustic", cp), (A.MAP_INDIRECT, "indirect", ip)):
        if len(g[key + "_pos"]) == 0:
            continue
        pv.select_map(which, float(np.sqrt(md2)), nlookup)
        Lr, Lt = pv.SurfaceLPhoton(g["sq_pts"], g["sq_n"], nlookup, md2, npaths)
        oLr, oLt = O.surface_lphoton(g[key + "_pos"], g[key + "_wi"], g[key + "_alpha"], g["sq_pts"], g["sq_n"], nlookup, md2, npaths)
        # exact ties at the n_lookup-th distance (coincident monochromatic photons) make the reference's own result order dependent
        tie = np.zeros(len(g["sq_pts"]), bool)
        t = O.KdTree(g[key + "_pos"])
        nf, idx, d2, _ = t.knn(g["sq_pts"], nlookup + 1, md2)
        full = nf > nlookup
        tie[full] = d2[full, nlookup] == d2[full, nlookup - 1]
        u = ~tie
        assert tie.mean() < 0.2
        for got, ref in ((Lr * inv_pi, g["slp_%s_Lr_pi" % key]), (Lt * inv_pi, g["slp_%s_Lt_pi" % key])):
            m = ref[u] > 0
            assert m.any()
            assert relerr(got[u], ref[u])[m].max() < 1e-4, key
            assert np.array_equal(got[u] == 0, ref[u] == 0), key
        with pytest.raises(Exception):
            pv.LPhoton(g["sq_pts"], g["sq_n"])                     # the volume estimate refuses to run on a surface map
    all_Lo = pv.RadiancePhotons(nlookup, md2, path_counts=(dp, ip, cp))
    pv.select_map(A.MAP_RADIANCE, float(np.sqrt(md2)), nlookup)
    idx, Lo = pv.RadianceNearest(g["sq_pts"], g["sq_n"])
    ref_idx = g["radn_idx"]
    oidx, od2 = O.radiance_nearest(g["rad_pos"], g["rad_n"], g["sq_pts"], g["sq_n"])
    assert np.array_equal(idx, oidx)                               # the stated rule: nearest facing photon, ties by index
    d2 = ((g["rad_pos"][idx] - g["sq_pts"]).astype(np.float32) ** 2)
    same = idx == ref_idx
    assert same.mean() > 0.95
    assert np.array_equal(od2.view(np.uint32), g["radn_d2"].view(np.uint32))   # ... at the reference's nearest distance, bit for bit
    assert np.array_equal(Lo, all_Lo[idx])                         # its radiance: the Lo pv_radiance_photons computed for that photon
    ref_Lo = g["radn_Lo"]                                          # (vs the reference: 1e-4 except at EPhoton tie sites, see the test above)
    assert np.quantile(relerr(Lo[same], ref_Lo[same])[ref_Lo[same] > 0], 0.9) < 1e-4
    # far-away and degenerate queries: still the global nearest facing photon
    far = np.array([[50, 50, 50], [-30, 0, 0], [0, 0, 0]], np.float32); fn = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    fi, _ = pv.RadianceNearest(far, fn, want_Lo=False)
    oi, _ = O.radiance_nearest(g["rad_pos"], g["rad_n"], far, fn)
    assert np.array_equal(fi, oi)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SURF_SCENES + ["sphere_glass"])
def test_final_gather_rays_vs_oracle(golden, pv_factory, pkg, name):
    """pv_final_gather (trace + nearest facing radiance photon at the hit + transmittance along the ray) against the composition of
    the pinned oracle's pieces (bvh_intersect / make_isect are bit-exact vs the reference, pvo_radiance_nearest and the
    transmittance march are pinned above) on the same Philox offsets: same radiance photon per ray, Lindir within 1e-4."""
    A = pkg._abi
    g, scene = golden(name)
    rng = np.random.default_rng(3)
    if name in SURF_SCENES:
        rp_pos, rp_n, rp_rho = g["rad_pos"], g["rad_n"], g["rad_rho_r"]
        Lo = g["rad_Lo"]
        origins = g["sq_pts"][:200] + 1e-3 * g["sq_n"][:200]
    else:                                              # a scene with a sphere: radiance photons scattered over its surfaces
        hit = g["hit_prim"] != 0xFFFFFFFF
        hr = g["hit_rays"][hit][:300]
        rp_pos = (hr["o"] + hr["d"] * g["hit_t"][hit][:300, None]).astype(np.float32)
        rp_n = -hr["d"] / np.linalg.norm(hr["d"], axis=1, keepdims=True); rp_n = rp_n.astype(np.float32)
        rp_rho = np.ones((len(rp_pos), 30), np.float32); Lo = rng.random((len(rp_pos), 30)).astype(np.float32)
        origins = g["q_pts"][:200]
    n = 1500
    o = origins[rng.integers(0, len(origins), size=n)].astype(np.float32)
    z = rng.uniform(-1, 1, size=n); phi = rng.uniform(0, 2 * np.pi, size=n); rr = np.sqrt(1 - z * z)
    d = np.stack([rr * np.cos(phi), rr * np.sin(phi), z], axis=1).astype(np.float32)
    rays = pkg.sceneio.make_rays(o, d, 1e-3, np.inf)
    istep = 0.05
    pv = pv_factory(stepsize=istep, seed=29); pv.set_scene(scene)
    pv.set_map_photons(A.MAP_RADIANCE, rp_pos, rp_n, rp_rho)
    # Lo of the radiance photons: computed by pv_radiance_photons from injected maps where the golden has them, else zero maps + a fake Lo
    if name in SURF_SCENES:
        nshot, cp, ip, dp, vp = [int(x) for x in g["counts"]]
        for which, key in ((A.MAP_CAUSTIC, "caustic"), (A.MAP_INDIRECT, "indirect"), (A.MAP_DIRECT, "direct")):
            pv.set_map_photons(which, g[key + "_pos"], g[key + "_wi"], g[key + "_alpha"])
        Lo = pv.RadiancePhotons(int(g["params"][6]), float(g["params"][7]), path_counts=(dp, ip, cp))
    else:
        # direct map = the radiance sites themselves with alpha chosen so that E is finite; only the plumbing matters here
        pv.set_map_photons(A.MAP_DIRECT, rp_pos, rp_n, Lo)
        Lo = pv.RadiancePhotons(20, 0.5, path_counts=(1000, 0, 0))
    assert (Lo > 0).any()
    pv.select_map(A.MAP_RADIANCE, 0.25, 50)
    L, idx = pv.FinalGather(rays, index_base=1000)
    oL, oidx = O.final_gather(scene, rp_pos, rp_n, Lo, rays, 4.0 * istep, seed=29, index_base=1000)
    assert (oidx != 0xFFFFFFFF).sum() > 0.5 * n
    same = idx == oidx
    # different radiance photon only for coincident photons (exact ties)
    for q in np.nonzero(~same)[0]:
        assert idx[q] != 0xFFFFFFFF and oidx[q] != 0xFFFFFFFF and np.array_equal(rp_pos[idx[q]], rp_pos[oidx[q]]), q
    assert same.mean() > 0.97
    m = (oL > 0) & same[:, None]
    assert m.any()
    assert relerr(L, oL)[m].max() < 1e-4
    assert np.array_equal((L == 0)[same], (oL == 0)[same])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell_homog", "cornell_grid32", "prism_small"])
def test_gather_step_parallel_equals_ray_parallel(golden, pv_factory, pkg, name):
    """The two schedules of the gather -- one warp per ray (frames) and one warp per march step + recurrence pass (small batches
    of secondary rays) -- run the same functions on the same inputs: L and T must agree bit for bit."""
    A = pkg._abi
    g, scene = golden(name)
    nused, maxdist, stepsize = int(g["params"][0]), float(g["params"][1]), float(g["params"][2])
    pv = pv_factory(stepsize=stepsize, nused=nused, maxdist=maxdist, seed=11)
    pv.set_scene(scene)
    pv.set_photons(g["shot_pos"], g["shot_wi"], g["shot_alpha"]); pv.build()
    rays = g["li_rays"]
    L0, T0 = pv.Li(rays, ray_index_base=5, flags=A.GATHER_RAY_PARALLEL)
    L1, T1 = pv.Li(rays, ray_index_base=5, flags=A.GATHER_STEP_PARALLEL)
    assert (L0 > 0).any()
    assert np.array_equal(L0.view(np.uint32), L1.view(np.uint32)) and np.array_equal(T0.view(np.uint32), T1.view(np.uint32))
    # the cell-batched schedule (the default in the fixed-radius regime) sums each step's photons in photon order: same neighbour
    # sets, T bit-identical, L to rounding; forced and default agree bit for bit where the default picks it
    pv.gather_stats(reset=True)
    L2, T2 = pv.Li(rays, ray_index_base=5, flags=A.GATHER_CELL_BATCHED)
    s2 = pv.gather_stats(reset=True)
    L0b, _ = pv.Li(rays, ray_index_base=5, flags=A.GATHER_RAY_PARALLEL)
    s0 = pv.gather_stats(reset=True)
    assert (s0.lookups, s0.photons_found, s0.heap_lookups) == (s2.lookups, s2.photons_found, s2.heap_lookups)
    assert np.array_equal(T0.view(np.uint32), T2.view(np.uint32))
    assert np.array_equal(L0 == 0, L2 == 0)
    m = L0 > 0
    assert (np.abs(L2 - L0)[m] / L0[m]).max() < 1e-5
    L4, T4 = pv.Li(rays, ray_index_base=5)                    # the default picks a schedule from the radius and the ray count
    assert np.array_equal(L4.view(np.uint32), L2.view(np.uint32)) or np.array_equal(L4.view(np.uint32), L0.view(np.uint32))
    L3, T3 = pv.Li(rays[:1], ray_index_base=5, flags=A.GATHER_STEP_PARALLEL)          # one ray
    assert np.array_equal(L0[:1].view(np.uint32), L3.view(np.uint32))


def test_malformed_bvh_is_refused(golden, pv_factory, pkg):
    """pv_set_scene walks the flattened BVH once on the host: a child offset outside the node array, a leaf whose primitive range
    leaves the primitive table, a split axis > 2 or a node reached twice is PV_EINVAL, not an out-of-bounds read in a kernel."""
    import copy
    g, scene = golden("cornell_homog")
    pv = pv_factory()
    pv.set_scene(scene)                                        # the real one is fine
    nodes = np.frombuffer(np.ascontiguousarray(scene.nodes).tobytes(), dtype=np.uint8).reshape(scene.n_nodes, 32).copy()
    interior = [i for i in range(scene.n_nodes) if nodes[i, 28] == 0]
    leaves = [i for i in range(scene.n_nodes) if nodes[i, 28] != 0]
    assert interior and leaves

    def broken(edit):
        s = copy.copy(scene)
        n = nodes.copy(); edit(n)
        s.nodes = n.reshape(-1)
        with pytest.raises(pkg.PVError) as e:
            pv.set_scene(s)
        return str(e.value)

    def set_offset(n, i, v): n[i, 24:28] = np.frombuffer(np.uint32(v).tobytes(), dtype=np.uint8)
    assert "second child" in broken(lambda n: set_offset(n, interior[0], scene.n_nodes + 7))
    assert "second child" in broken(lambda n: set_offset(n, interior[0], interior[0]))           # a cycle
    assert "primitive range" in broken(lambda n: set_offset(n, leaves[0], scene.n_prims))
    assert "axis" in broken(lambda n: n.__setitem__((interior[0], 29), 3))
    pv.set_scene(scene)


def test_shooter_deep_continuation_stacks(golden, pv_factory, monkeypatch):
    """Paths that scatter again and again (Q1 inverted: a medium with sigma_a > sigma_s scatters MOST interactions, and every
    scatter leaves a continuation frame behind, Q2): the frames above the fourth level live in pages from a pool, and a wave that
    runs the pool dry is replayed with a larger one.  Started from a pool of ONE page the photon set is the same as with the
    default pool, and it is the oracle's."""
    import copy
    g, scene = golden("cornell_grid32")
    s = copy.copy(scene)
    s.medium = type(scene.medium).from_buffer_copy(scene.medium)          # (a ctypes struct with a pointer: copied byte for byte)
    for b in range(30):
        s.medium.sigma_a[b] = 3.0; s.medium.sigma_s[b] = 2.0
    istep = float(g["params"][2])
    sets = []
    for pages in (None, "1"):
        if pages is None: monkeypatch.delenv("PV_WF_DEEP_PAGES", raising=False)
        else: monkeypatch.setenv("PV_WF_DEEP_PAGES", pages)
        pv = pv_factory(stepsize=istep, seed=5)
        pv.set_scene(s)
        st = pv.Preprocess(20000, stepsize=0.05, max_photon_depth=5, build=False)
        assert st.stack_overflows == 0
        sets.append(pv.get_photons())
    for a, b in zip(*sets):
        assert np.array_equal(a, b)
    ref = O.shoot(s, 20000, 0.05, istep, seed=5, rng_mode=O.PHILOX, nthreads=8)
    ids = sets[0][3]
    common = np.intersect1d(ids, ref["ids"])
    assert len(common) >= 0.995 * max(len(ids), len(ref["ids"]))
    # deposit ordinals beyond 4 on one path = at least that many scatters deep
    assert (ids & 0xffff).max() >= 5
