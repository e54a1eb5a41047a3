"""Multi-GPU partitioning of the `generate` workload.

Every clip is independent end to end (per-clip peak normalisation flowhighsr.py:69, per-clip
attention, per-clip post-processing cutoff and normalisation postprocessing.py:26,40), so the path
shards by clip with NO data-path collective: one process per GPU, weights replicated, rank r takes
the clips `assign_clips` gives it.  A collective appears only when the caller wants all outputs
on one rank (`gather_outputs`: torch.distributed all_gather_object of fp32 waveforms).

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
