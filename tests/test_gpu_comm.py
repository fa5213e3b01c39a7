"""GPU (-m gpu): NCCL behind the C ABI (csrc/pv_comm.cu) with a one-rank communicator -- the call sequence every rank of a
multi-GPU run makes (pv_comm_unique_id -> pv_comm_init -> pv_allgather_photons), checked for what it must leave behind.
The N > 1 exchange itself runs in bench.py under torchrun (SCALE_r*.json); its host logic is covered by the gloo tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_allgather_of_one_rank_is_the_identity(golden, pv_factory, pkg):
    g, scene = golden("cornell_homog")
    pv = pv_factory(stepsize=0.05, nused=50, maxdist=0.25, seed=3)
    pv.set_scene(scene)
    pos, wi, alpha = g["shot_pos"], g["shot_wi"], g["shot_alpha"]
    pv.set_photons(pos, wi, alpha)
    pv.comm_init(pkg.PhotonVolume.comm_unique_id(), 0, 1)
    ms = pv.allgather_photons(renumber=True)
    assert ms >= 0.0
    p2, w2, a2, ids = pv.get_photons()
    assert np.array_equal(ids, np.arange(len(pos), dtype=np.uint64))
    assert np.array_equal(p2, pos) and np.array_equal(w2, wi) and np.array_equal(a2, alpha)
    # the map builds and answers as before
    pv.build()
    nf, idx, d2 = pv.Lookup(g["q_pts"][:64], k=50, r2=0.25 ** 2)
    assert np.array_equal(idx, g["knn50_idx"][:64])
    pv.comm_destroy()


def test_allgather_keeps_a_shot_photon_set(golden, pv_factory, pkg):
    g, scene = golden("cornell_homog")
    pv = pv_factory(stepsize=0.05, seed=9)
    pv.set_scene(scene)
    pv.Preprocess(2000, stepsize=0.05, build=False)
    before = pv.get_photons()
    pv.comm_init(pkg.PhotonVolume.comm_unique_id(), 0, 1)
    pv.allgather_photons()
    after = pv.get_photons()
    for a, b in zip(before, after):
        assert np.array_equal(a, b)


def test_allgather_without_a_communicator_is_an_error(golden, pv_factory, pkg):
    g, scene = golden("cornell_homog")
    pv = pv_factory(stepsize=0.05, seed=9)
    pv.set_scene(scene)
    with pytest.raises(pkg.PVError) as e:
        pv.allgather_photons()
    assert "communicator" in str(e.value)
