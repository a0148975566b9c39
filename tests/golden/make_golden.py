"""Generates tests/golden/fm_golden.npz and tests/golden/tiny_ref.fmi from the REFERENCE ITSELF
(oracle/_ref/libseal_ref.so = unmodified seal/cpp_modules/fm_index.cpp + vendored sdsl-lite).
Run in the build container only (needs /root/reference for keeper.int and the compiled _ref):

    python tests/golden/make_golden.py

Fixtures:
  keeper     sdsl-lite's own vendored test input test/test_cases/keeper.int (63 u64); symbols
             shifted by +1 so that 0 stays the sentinel.
  toy        the vector of the commented-out main() in fm_index.cpp:203 (+10 shift).
  rand5k     5 000 random symbols over 300 values (seed 11).
  phrase     400 docs x 25 tokens of the phrase generator (seed 5), SEAL layout (reversed, +10).
For each: first-step (sym, 0, size()) results for every symbol (pins SURVEY.md §H1), 400 random
8-step walks with every intermediate (lo,hi), distinct_count of every visited range, locate of 300
rows, extract_text of 100 intervals, plus the raw sections of the saved .fmi for `phrase`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.fm_oracle import RefFM  # noqa: E402
from seal_b200.synthetic import make_corpus, corpus_symbols  # noqa: E402


def record(name, text, out, rng):
    r = RefFM(text)
    size = r.size()
    out[f"{name}.text"] = np.asarray(text, dtype=np.uint64)
    out[f"{name}.size"] = np.uint64(size)
    smax = int(text.max()) + 12
    syms = np.arange(0, smax, dtype=np.uint64)
    fl, fh = r.backward_search_step_batch(syms, np.zeros_like(syms), np.full_like(syms, size))
    out[f"{name}.first_lo"], out[f"{name}.first_hi"] = fl, fh
    W, D = 400, 8
    wsym = np.zeros((W, D), dtype=np.uint64); wlo = np.zeros((W, D), dtype=np.uint64); whi = np.zeros((W, D), dtype=np.uint64)
    dc_off = [0]; dc = []
    for w in range(W):
        lo, hi = 0, size
        for d in range(D):
            s = int(text[rng.integers(0, len(text))]) if rng.random() < 0.93 else int(rng.integers(0, smax))
            lo, hi = r.backward_search_step(s, lo, hi)
            wsym[w, d], wlo[w, d], whi[w, d] = s, lo, hi
            v = r.distinct_count(lo, hi + 1) if hi + 1 >= lo else np.zeros(0, dtype=np.uint64)
            dc.append(v); dc_off.append(dc_off[-1] + len(v))
    out[f"{name}.walk_sym"], out[f"{name}.walk_lo"], out[f"{name}.walk_hi"] = wsym, wlo, whi
    out[f"{name}.dc_off"] = np.asarray(dc_off, dtype=np.uint64)
    out[f"{name}.dc"] = np.concatenate(dc) if dc else np.zeros(0, dtype=np.uint64)
    rows = rng.integers(0, size + 2, size=300).astype(np.uint64)
    out[f"{name}.loc_rows"] = rows
    out[f"{name}.loc"] = np.asarray([r.locate(int(x)) for x in rows], dtype=np.uint64)
    n = len(text)
    b = rng.integers(0, n, size=100); e = np.minimum(b + rng.integers(0, 30, size=100), n)
    out[f"{name}.ext_b"], out[f"{name}.ext_e"] = b.astype(np.uint64), e.astype(np.uint64)
    ex = [r.extract_text(int(x), int(y)) for x, y in zip(b, e)]
    out[f"{name}.ext_off"] = np.asarray(np.cumsum([0] + [len(x) for x in ex]), dtype=np.uint64)
    out[f"{name}.ext"] = np.concatenate(ex) if ex else np.zeros(0, dtype=np.uint64)
    return r


def main():
    rng = np.random.default_rng(2024)
    out = {}
    keeper = np.fromfile("/root/reference/res/external/sdsl-lite/test/test_cases/keeper.int", dtype=np.uint64)
    record("keeper", keeper + 1, out, rng)
    record("toy", np.array([1, 8, 15, 23, 1, 8, 23, 11, 8], dtype=np.uint64) + 10, out, rng)
    record("rand5k", np.random.default_rng(11).integers(10, 310, size=5000).astype(np.uint64), out, rng)
    docs = make_corpus(n_docs=400, doc_len=25, n_phrases=800, seed=5)
    out["phrase.docs"] = docs
    r = record("phrase", corpus_symbols(docs), out, rng)
    r.save(os.path.join(HERE, "tiny_ref.fmi"))
    np.savez_compressed(os.path.join(HERE, "fm_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "fm_golden.npz"), os.path.getsize(os.path.join(HERE, "fm_golden.npz")), "bytes;",
          "tiny_ref.fmi", os.path.getsize(os.path.join(HERE, "tiny_ref.fmi")), "bytes")


if __name__ == "__main__":
    main()
