"""CPU, world_size 2, gloo: the N>1 path's host logic — block sharding of the query batch and the
single gather of fixed-size hypothesis records — gives exactly the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_generate(ids, mask, H=7, T=5):
    """Deterministic stand-in for generate_records: records are a pure function of each query."""
    ids = np.asarray(ids); Q = ids.shape[0]
    base = (ids * np.asarray(mask)).sum(axis=1).astype(np.int64)
    scores = (base[:, None] * 0.001 - np.arange(H)[None, :]).astype(np.float32)
    lens = ((base[:, None] + np.arange(H)[None, :]) % T + 1).astype(np.int32)
    toks = ((base[:, None, None] + np.arange(H)[None, :, None] * 3 + np.arange(T)[None, None, :]) % 1000).astype(np.int32)
    valid = ((base[:, None] + np.arange(H)[None, :]) % 3).astype(np.uint8)
    lo = (base[:, None] * 7 + np.arange(H)[None, :]).astype(np.uint64)
    return {"scores": scores, "lens": lens, "tokens": toks, "valid": valid, "lo": lo, "hi": lo + 5}


def _worker(rank, world, port, n_queries, q):
    sys.path.insert(0, ROOT)
    from seal_b200.sharding import sharded_generate, pack_host_records
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    ids = rng.integers(0, 1000, size=(n_queries, 9)); mask = (rng.random((n_queries, 9)) < 0.8).astype(np.int64)

    def fill(ids_blk, mask_blk, layout):            # what the decode kernels do on the device, done on the host here
        rec = fake_generate(ids_blk, mask_blk) if len(ids_blk) else {}
        return torch.from_numpy(pack_host_records(rec, layout, len(ids_blk)))

    out = sharded_generate(fill, ids, mask, hyps=7, max_length=5)
    if rank == 0:
        full = fake_generate(ids, mask)
        ok = all(np.array_equal(out[k], full[k]) for k in full)
        q.put(bool(ok) and out["scores"].shape[0] == n_queries and not out["errors"].any())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("n_queries", [10, 7, 1])
def test_sharded_generate_equals_unsharded_world2(n_queries):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_queries, q)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True


def test_shard_bounds_cover_everything():
    from seal_b200.sharding import shard_bounds
    for n in (0, 1, 7, 1000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_record_layout_round_trip():
    from seal_b200.sharding import RecordLayout, pack_host_records, merge_gathered
    lay = RecordLayout(5, 7, 5)
    assert lay.nbytes % 16 == 0 and all(off % 16 == 0 for off, *_ in lay.offsets.values())
    assert lay.record_bytes == 5 * 7 * (8 + 8 + 4 + 4 + 4 * 5 + 1)
    rng = np.random.default_rng(1)
    ids = rng.integers(0, 1000, size=(3, 9)); mask = np.ones_like(ids)
    rec = fake_generate(ids, mask)
    back = merge_gathered([pack_host_records(rec, lay, 3)], lay)
    assert all(np.array_equal(back[k], rec[k]) for k in rec) and back["scores"].shape[0] == 3
