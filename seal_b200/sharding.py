"""Multi-GPU: query sharding with replicated index + weights and ONE gather of result records.

Queries are independent (own beams, SA ranges, hypotheses); the FM-index and BART weights are
read-only, so every rank (one process per GPU) decodes a contiguous block of the batch and the
fixed-size hypothesis records are gathered once to rank 0 — the only collective.
Works with any torch.distributed backend: NCCL on the B200s, gloo in the CPU tests.
"""
from typing import Callable, Dict, Optional

import numpy as np


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one query."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(rec: Dict[str, np.ndarray], n_local: int, n_max: int, device=None, dst: int = 0, group=None):
    """Gathers per-query record arrays (first dim = queries) to `dst` with a single collective:
    every field is padded to n_max queries, viewed as bytes and packed into one uint8 buffer."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    keys = sorted(k for k, v in rec.items() if v is not None)
    parts, meta = [], []
    for k in keys:
        a = np.ascontiguousarray(rec[k])
        pad = np.zeros((n_max,) + a.shape[1:], dtype=a.dtype)
        pad[:n_local] = a[:n_local]
        b = pad.view(np.uint8).reshape(-1)
        meta.append((k, a.dtype, a.shape[1:], b.size))
        parts.append(b)
    header = np.asarray([n_local], dtype=np.int64).view(np.uint8)
    buf = torch.from_numpy(np.concatenate([header] + parts))
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)                 # the one collective
    if rank != dst:
        return None
    merged = {k: [] for k in keys}
    for t in out:
        raw = t.cpu().numpy()
        n = int(raw[:8].view(np.int64)[0]); off = 8
        for k, dt, shp, size in meta:
            arr = raw[off:off + size].view(dt).reshape((n_max,) + tuple(shp)); off += size
            merged[k].append(arr[:n])
    return {k: np.concatenate(v, axis=0) for k, v in merged.items()}


def sharded_generate(generate_fn: Callable[..., Dict[str, np.ndarray]], input_ids, attention_mask, device=None,
                     group=None, **kw) -> Optional[Dict[str, np.ndarray]]:
    """Runs `generate_fn(input_ids[lo:hi], attention_mask[lo:hi], **kw)` (e.g. a closure over
    seal_b200.beam_search.generate_records) on this rank's block and gathers the records to rank 0.
    Returns the full-batch records on rank 0, None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    n = len(input_ids)
    lo, hi = shard_bounds(n, world, rank)
    n_max = max(shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world))
    if hi > lo:
        rec = generate_fn(input_ids[lo:hi], attention_mask[lo:hi], **kw)
    else:
        rec = None
    if rec is None:                                             # empty shard: learn the layout from a neighbour-free dummy
        rec = generate_fn(input_ids[:1], attention_mask[:1], **kw)
        rec = {k: (None if v is None else v[:0]) for k, v in rec.items()}
    return gather_records(rec, hi - lo, n_max, device=device, group=group)
