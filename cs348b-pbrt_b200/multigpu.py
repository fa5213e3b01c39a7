"""One-process-per-GPU plumbing of the path: emission sharded by 4096-path blocks, ONE all-gather that
replicates the photon map over NVLink, image tiles dealt round-robin.  torch.distributed is only the transport; the
functions work on any backend/device (NCCL + CUDA in bench.py, gloo + CPU in tests/test_multigpu_host.py)."""
import numpy as np


def last_block(counts, first_block, running_total, target):
    """PhotonShootingTask::Run's stop rule (photonshooter.cpp:330-340) on globally summed per-block deposit counts:
    returns (last_block or 0, new running total, blocks consumed).  Blocks are 1-based."""
    total = int(running_total)
    for i, c in enumerate(counts):
        total += int(c)
        if total >= target:
            return first_block + i, total, i + 1
    return 0, total, len(counts)


def next_wave(total, blocks_done, target, world, cap=262144):
    """Number of blocks to trace next, sized from the observed deposit yield (a little past the target)."""
    per_block = max(total / max(blocks_done, 1), 1e-3)
    return int(min(max((target - total) / per_block * 1.03 + 8, 64 * world), cap))


def allreduce_counts(dist, torch, device):
    """The callback PhotonVolume.PreprocessMapsRanks wants: sums a uint32 numpy array in place over all ranks (per-class,
    per-block deposit counts of one wave of the sharded all-maps pass)."""
    def fn(arr):
        t = torch.from_numpy(arr.astype(np.int64)).to(device)
        dist.all_reduce(t)
        arr[:] = t.cpu().numpy().astype(np.uint32)
    return fn


def merge_by_id(parts):
    """Per-rank photon lists (pos, wi, alpha, ids), each ordered by id -> the global list ordered by id (= the single-rank order)."""
    ids = np.concatenate([p[3] for p in parts])
    order = np.argsort(ids, kind="stable")
    return tuple(np.concatenate([p[k] for p in parts])[order] for k in range(4))


def allgather_photons(dist, torch, pos, wi, alpha, device):
    """Replicate the photon planes of all ranks (rank order, i.e. global photon order).  Slices may differ in length:
    planes are padded to the longest slice for all_gather_into_tensor and the padding is dropped afterwards.
    Returns (pos, wi, alpha, n_total, seconds_of_the_collective or None)."""
    world = dist.get_world_size()
    n_local = int(len(pos))
    cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    outs = []
    for plane, k in ((pos, 3), (wi, 3), (alpha, 30)):
        loc = torch.zeros((mx, k), dtype=torch.float32, device=device)
        if n_local:
            loc[:n_local].copy_(plane if torch.is_tensor(plane) else torch.from_numpy(np.ascontiguousarray(plane, dtype=np.float32)))
        full = torch.empty((world * mx, k), dtype=torch.float32, device=device)
        outs.append((loc, full))
    seconds = None
    if device.type == "cuda":
        torch.cuda.synchronize(device)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    for loc, full in outs:
        dist.all_gather_into_tensor(full, loc)
    if device.type == "cuda":
        e1.record(); torch.cuda.synchronize(device)
        seconds = e0.elapsed_time(e1) * 1e-3
    keep = torch.cat([torch.arange(r * mx, r * mx + counts[r], device=device) for r in range(world)])
    planes = [full.index_select(0, keep).contiguous() for _, full in outs]
    return planes[0], planes[1], planes[2], sum(counts), seconds
