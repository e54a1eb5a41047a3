"""Multi-GPU partitioning of the `generate` workload.

Every clip is independent end to end (per-clip peak normalisation flowhighsr.py:69, per-clip
attention, per-clip post-processing cutoff and normalisation postprocessing.py:26,40), so the path
shards by clip with NO data-path collective: one process per GPU, weights replicated, rank r takes
the clips `assign_clips` gives it.  A collective appears only when the caller wants all outputs
on every rank: `gather_clip_tensor` / `gather_chunk_tensor` = ONE `all_gather_into_tensor` (ncclAllGather over
NVLink / NVSwitch, or gloo on CPU) of the fp32 waveforms, padded to equal per-rank blocks.
(`gather_outputs` is the ragged, pickling fallback for clips of unequal length.)

Long-form audio (BASELINE config 4) is cut into overlapped chunks in the 48 kHz domain; every chunk
runs the per-clip pipeline and the chunks are stitched by a linear cross-fade overlap-add.
Chunking is a capability of this engine, not of the reference (whose O(N^2) fp32 attention cannot
run N = 60 000 frames).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def assign_clips(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first balance of clip indices over ranks (equal lengths -> round robin)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    for r in range(world_size):
        out[r].sort()
    return out


def chunk_plan(total: int, chunk: int, overlap: int) -> List[Tuple[int, int]]:
    """[start, end) spans of overlapped chunks covering [0, total); consecutive spans share `overlap`."""
    if chunk <= overlap:
        raise ValueError("chunk must be longer than overlap")
    if total <= chunk:
        return [(0, total)]
    step = chunk - overlap
    spans = []
    s = 0
    while True:
        e = min(s + chunk, total)
        spans.append((s, e))
        if e == total:
            break
        s += step
    if len(spans) > 1 and spans[-1][1] - spans[-1][0] <= overlap:  # merge a tiny tail chunk
        last = spans.pop()
        spans[-1] = (spans[-1][0], last[1])
    return spans


def overlap_add(chunks: Sequence[np.ndarray], spans: Sequence[Tuple[int, int]], total: int) -> np.ndarray:
    """Linear cross-fade stitch: inside an overlap the weights of the two chunks sum to 1."""
    out = np.zeros(total, dtype=np.float64)
    wsum = np.zeros(total, dtype=np.float64)
    for k, (c, (s, e)) in enumerate(zip(chunks, spans)):
        n = e - s
        assert c.shape[-1] == n, (c.shape, n)
        w = np.ones(n, dtype=np.float64)
        if k > 0:
            ov = spans[k - 1][1] - s
            if ov > 0:
                w[:ov] = (np.arange(ov) + 0.5) / ov
        if k + 1 < len(spans):
            ov = e - spans[k + 1][0]
            if ov > 0:
                w[n - ov:] = 1.0 - (np.arange(ov) + 0.5) / ov
        out[s:e] += w * c
        wsum[s:e] += w
    return (out / np.maximum(wsum, 1e-12)).astype(np.float32)


def block_range(n: int, world_size: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous block partition of n items: -> (start, end, per_rank) with per_rank = ceil(n / world_size);
    ranks past the end get empty blocks.  Used for long-form chunks (neighbouring chunks stay on one GPU)."""
    per = -(-n // world_size) if n > 0 else 0
    s = min(rank * per, n)
    return s, min(s + per, n), per


def gather_blocks(local, per_rank: int, n: int, group=None):
    """local: torch tensor [<= per_rank, ...] (this rank's block of `block_range`) -> [n, ...] on every rank with ONE
    all_gather_into_tensor (ncclAllGather on CUDA tensors, gloo on CPU tensors).  Blocks are zero-padded to per_rank rows."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:n]
    pad = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n]


def gather_assigned(local, mine: Sequence[int], parts: Sequence[Sequence[int]], group=None):
    """Clips distributed with `assign_clips` (all of one length): local [len(mine), T] -> [n_clips, T] in clip order on
    every rank, one all_gather_into_tensor + an index permutation."""
    import torch
    n = sum(len(p) for p in parts)
    per = max(len(p) for p in parts) if parts else 0
    flat = gather_blocks(local, per, len(parts) * per, group) if len(parts) > 1 else local
    if len(parts) == 1:
        return local
    out = torch.empty((n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, p in enumerate(parts):
        if len(p):
            idx = torch.as_tensor(list(p), device=local.device)
            out[idx] = flat[r * per: r * per + len(p)]
    return out


def gather_outputs(local: Dict[int, np.ndarray], world_size: int, rank: int, group=None) -> Dict[int, np.ndarray]:
    """Collects {clip index -> waveform} from every rank on every rank (nccl or gloo)."""
    if world_size == 1:
        return dict(local)
    import torch.distributed as dist
    bucket: List[Dict[int, np.ndarray]] = [None] * world_size  # type: ignore[list-item]
    dist.all_gather_object(bucket, local, group=group)
    merged: Dict[int, np.ndarray] = {}
    for d in bucket:
        merged.update(d)
    return merged
