"""One large batch of LF steps (lf_step_kernel) on an index of `n_tokens` tokens, for an `ncu --set full` capture:
   ncu --set full --clock-control none -k regex:lf_step --launch-skip 2 -c 1 -o out python tools/lf_ncu_probe.py 200000000
Prints the CUDA-event time of the batch as well (not valid under ncu)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seal_b200.cpp_modules.fm_index import FMIndex  # noqa: E402
from seal_b200.synthetic import make_corpus, corpus_symbols  # noqa: E402

n_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
rng = np.random.default_rng(7)
if n_tokens <= 10_000_000:
    text = corpus_symbols(make_corpus())
else:                                                   # i.i.d. Zipf over 50 k symbols (tools/fm_microbench.py FMB_BIG)
    p = 1.0 / (np.arange(50000) + 1.0); cdf = np.cumsum(p / p.sum())
    text = (np.minimum(np.searchsorted(cdf, rng.random(n_tokens)), 49999) + 14).astype(np.uint64)
fm = FMIndex(); fm.initialize(text); fm.to_device(0)
m = fm.size()
sym = torch.tensor(text[rng.integers(0, len(text), size=N)].astype(np.int64), device="cuda")
lo = torch.zeros(N, dtype=torch.int64, device="cuda"); hi = torch.full((N,), m - 1, dtype=torch.int64, device="cuda")
lo1, hi1 = fm.lf_step_tensors(sym, lo, hi)              # depth 1: wide ranges
sym2 = torch.tensor(text[rng.integers(0, len(text), size=N)].astype(np.int64), device="cuda")
for _ in range(3):
    fm.lf_step_tensors(sym2, lo1, hi1)                  # depth 2 from data-dependent (lo, hi): the measured launch
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); fm.lf_step_tensors(sym2, lo1, hi1); e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3
print(f"n_tokens {n_tokens} size {m}: {N} LF steps in {us:.1f} us = {N / us / 1e3:.2f} G steps/s, algorithmic {N * 768 / us / 1e3:.0f} GB/s")
