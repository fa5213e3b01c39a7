"""CPU: the host-side logic of the N > 1 path with world_size 2 over gloo (no GPU): block dealing and the stop rule of the
sharded photon pass, the photon all-gather with ragged slices, and the tile sharding of the frame."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from __graft_entry__ import load_package

pkg = load_package()
from cs348b_pbrt_b200 import multigpu as MG, workloads as W  # noqa: E402


def test_last_block_rule():
    assert MG.last_block([3, 4, 5], 1, 0, 7) == (2, 7, 2)          # reached exactly at block 2
    assert MG.last_block([3, 4, 5], 10, 0, 100) == (0, 12, 3)      # not reached: continue
    assert MG.last_block([0, 0, 9], 5, 1, 10) == (7, 10, 3)
    assert MG.next_wave(100, 10, 1000, 2) >= 90


def test_merge_by_id_restores_the_single_rank_order():
    rng = np.random.default_rng(1)
    ids = np.sort(rng.choice(10_000, size=500, replace=False)).astype(np.uint64)
    pos = rng.random((500, 3)).astype(np.float32); wi = rng.random((500, 3)).astype(np.float32); alpha = rng.random((500, 30)).astype(np.float32)
    pick = rng.random(500) < 0.4                                   # photons of "rank 0"; both parts stay ordered by id
    parts = [(pos[m], wi[m], alpha[m], ids[m]) for m in (pick, ~pick)]
    got = MG.merge_by_id(parts)
    assert np.array_equal(got[3], ids) and np.array_equal(got[0], pos) and np.array_equal(got[2], alpha)


def test_photon_slices_partition_the_set():
    n = 7_300_000
    for world in (1, 2, 3, 8):
        edges = [W.photon_slice(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_tiles_cover_the_frame_once():
    cfg = dict(W.CONFIGS["tiny"])
    seen = np.zeros(cfg["xres"] * cfg["yres"], np.int32)
    for r in range(3):
        rays, order = W.frame_rays(cfg, r, 3)
        assert len(rays) == len(order)
        seen[order] += 1
    assert (seen == 1).all()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    # ragged slices of one global photon set
    n = 1000
    rng = np.random.default_rng(0)
    pos = rng.random((n, 3)).astype(np.float32); wi = rng.random((n, 3)).astype(np.float32); alpha = rng.random((n, 30)).astype(np.float32)
    lo, hi = (0, 377) if rank == 0 else (377, n)
    p, w, a, n_all, _ = MG.allgather_photons(dist, torch, pos[lo:hi], wi[lo:hi], alpha[lo:hi], dev)
    ok = n_all == n and np.array_equal(p.numpy(), pos) and np.array_equal(w.numpy(), wi) and np.array_equal(a.numpy(), alpha)
    # blocks are dealt (b - 1) % world == rank; per-block counts are all-reduced, every rank must agree on the last block
    counts = np.zeros(16, np.int64)
    for i in range(16):
        if i % world == rank:
            counts[i] = 10 + i
    t = torch.from_numpy(counts.copy()); dist.all_reduce(t)
    last, total, used = MG.last_block(t.numpy(), 1, 0, 100)
    # the all-reduce callback of the sharded all-maps pass (uint32 counts in place) and the id merge of per-rank photon lists
    cnt = np.arange(12, dtype=np.uint32) * (rank + 1)
    MG.allreduce_counts(dist, torch, dev)(cnt)
    ok = ok and np.array_equal(cnt, np.arange(12, dtype=np.uint32) * 3)
    q.put((rank, ok, last, total, used))
    dist.destroy_process_group()


def test_world2_gloo_allgather_and_stop_rule():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert res[0][2:] == res[1][2:]
    expect = MG.last_block([10 + i for i in range(16)], 1, 0, 100)
    assert res[0][2:] == expect
