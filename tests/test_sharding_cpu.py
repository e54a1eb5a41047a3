"""Host-side multi-GPU logic on CPU: clip assignment, chunk plan / overlap-add, and a world_size-2
gloo run of the gather (the data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowhigh_b200 import sharding


def test_assign_clips_balanced_and_complete():
    lengths = [480000] * 512
    parts = sharding.assign_clips(lengths, 8)
    assert sorted(i for p in parts for i in p) == list(range(512))
    assert all(len(p) == 64 for p in parts)
    ragged = [80000, 120000, 160000, 240000] * 5 + [10]
    parts = sharding.assign_clips(ragged, 4)
    loads = [sum(ragged[i] for i in p) for p in parts]
    assert sorted(i for p in parts for i in p) == list(range(len(ragged)))
    assert max(loads) - min(loads) <= 240000
    assert sharding.assign_clips([], 2) == [[], []]


@pytest.mark.parametrize("total,chunk,overlap", [(28_800_000, 480_000, 24_000), (1000, 300, 50), (250, 300, 50), (640, 300, 50)])
def test_chunk_plan_and_overlap_add_reconstruct(total, chunk, overlap):
    spans = sharding.chunk_plan(total, chunk, overlap)
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
        assert s0 < s1 < e0  # consecutive chunks overlap
    rng = np.random.default_rng(0)
    x = rng.standard_normal(total).astype(np.float32)
    y = sharding.overlap_add([x[s:e] for s, e in spans], spans, total)
    assert np.abs(y - x).max() <= 1e-6  # stitching identical chunks is the identity


def test_block_range_covers_everything():
    for n, w in ((64, 8), (5, 2), (3, 4), (0, 2), (61, 8)):
        spans = [sharding.block_range(n, w, r) for r in range(w)]
        assert [i for s, e, _ in spans for i in range(s, e)] == list(range(n))
        assert len({p for _, _, p in spans}) == 1 and all(e - s <= p for s, e, p in spans)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [100 + i for i in range(7)]
    mine = sharding.assign_clips(lengths, world)[rank]
    local = {i: np.full(lengths[i], float(i), np.float32) for i in mine}
    merged = sharding.gather_outputs(local, world, rank)
    ok = sorted(merged) == list(range(7)) and all(merged[i].shape[0] == lengths[i] and merged[i][0] == i for i in merged)
    # tensor gathers (one all_gather_into_tensor each): contiguous blocks (long-form chunks) and assigned clips
    K = 5
    k0, k1, per = sharding.block_range(K, world, rank)
    blk = torch.arange(k0, k1, dtype=torch.float32)[:, None].repeat(1, 3)
    allc = sharding.gather_blocks(blk, per, K)
    ok = ok and allc.shape == (K, 3) and torch.equal(allc[:, 0], torch.arange(K, dtype=torch.float32))
    parts = sharding.assign_clips([480] * 7, world)
    loc = torch.tensor([float(i) for i in parts[rank]])[:, None].repeat(1, 4)
    allp = sharding.gather_assigned(loc, parts[rank], parts)
    ok = ok and allp.shape == (7, 4) and torch.equal(allp[:, 1], torch.arange(7, dtype=torch.float32))
    t = torch.tensor([float(rank + 1)])  # bench.py reduces its timing the same way: max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, bool(ok and float(t) == world)))
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
