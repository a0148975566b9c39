"""Runs the REFERENCE'S OWN seal/index.py (imported unmodified) on top of the compiled reference
fm_index.cpp (oracle/_ref, reached through a SWIG-shadow-shaped ctypes class, since SWIG is not installed),
checks oracle.fm_oracle.OracleIndex against every method, and stores the reference's answers in
tests/golden/index_golden.json.  Run in the build container only:

    python tests/golden/make_index_golden.py
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.fm_oracle import RefFM, OracleIndex  # noqa: E402
from seal_b200.synthetic import make_corpus  # noqa: E402


class ShadowFMIndex:
    """Shape of the SWIG shadow class of fm_index.hpp:20-45 (methods return Python lists of ints)."""
    def __init__(self): self._r = None
    def initialize(self, data): self._r = RefFM(np.asarray(data, dtype=np.uint64))
    def initialize_from_file(self, file, width): self._r = RefFM(from_file=(file, width))
    def size(self): return self._r.size()
    def backward_search_multi(self, q): return list(self._r.backward_search_multi(q))
    def backward_search_step(self, s, lo, hi): return list(self._r.backward_search_step(s, lo, hi))
    def distinct(self, lo, hi): return self._r.distinct(lo, hi).tolist()
    def distinct_count(self, lo, hi): return self._r.distinct_count(lo, hi).tolist()
    def distinct_count_multi(self, lows, highs): return [x.tolist() for x in self._r.distinct_count_multi(lows, highs)]
    def locate(self, row): return self._r.locate(row)
    def extract_text(self, b, e): return self._r.extract_text(b, e).tolist()
    def save(self, path): self._r.save(path)


def load_FMIndex(path):
    x = ShadowFMIndex(); x._r = RefFM(path=path)
    return x


def load_reference_index():
    seal = types.ModuleType("seal"); seal.__path__ = []
    cpp = types.ModuleType("seal.cpp_modules"); cpp.__path__ = []
    fm = types.ModuleType("seal.cpp_modules.fm_index"); fm.FMIndex = ShadowFMIndex; fm.load_FMIndex = load_FMIndex
    sys.modules.update({"seal": seal, "seal.cpp_modules": cpp, "seal.cpp_modules.fm_index": fm})
    spec = importlib.util.spec_from_file_location("seal.index", "/root/reference/seal/index.py")
    mod = importlib.util.module_from_spec(spec); sys.modules["seal.index"] = mod; spec.loader.exec_module(mod)
    return mod


CORPUS = dict(n_docs=200, doc_len=24, n_phrases=300, seed=21, vocab=500)


def probe(ix, docs, rng_seed=5):
    """Every public method of seal/index.py on a fixed set of arguments -> JSON-able dict."""
    rng = np.random.default_rng(rng_seed)
    out = {"len": len(ix), "n_docs": ix.n_docs, "size": ix.size(), "beginnings_tail": list(ix.beginnings[-3:]),
           "occurring_distinct": list(ix.occurring_distinct), "occurring_counts": list(ix.occurring_counts)}
    seqs = [[]]
    for _ in range(60):
        d = int(rng.integers(0, docs.shape[0])); L = int(rng.integers(1, 6)); a = int(rng.integers(0, docs.shape[1] - L))
        s = [int(t) for t in docs[d, a:a + L]]
        if rng.random() < 0.15:
            s[-1] = int(rng.integers(4, 500))
        seqs.append(s)
    seqs += [[int(t)] for t in range(0, 40)]                       # every small id as a unigram (covers the H1 first step)
    out["seqs"] = seqs
    out["get_range"] = [list(ix.get_range(s)) for s in seqs]
    out["get_count"] = [ix.get_count(s) for s in seqs]
    out["get_continuations"] = [list(ix.get_continuations(s)) for s in seqs[:30]]
    out["get_doc_indices"] = [list(ix.get_doc_indices(s))[:50] for s in seqs[1:30]]
    rr = [r for r in out["get_range"] if r[1] > r[0]][:40]
    out["ranges"] = rr
    out["get_distinct"] = [list(ix.get_distinct(a, b)) for a, b in rr]
    out["get_distinct_count"] = [[list(x) for x in ix.get_distinct_count(a, b)] for a, b in rr]
    out["get_distinct_count_multi"] = [[list(x) for x in p] for p in ix.get_distinct_count_multi([a for a, _ in rr], [b for _, b in rr])]
    rows = [int(x) for x in rng.integers(0, ix.size(), size=60)] + [0, ix.size() - 1, ix.size(), ix.size() + 5]
    out["rows"] = rows
    out["get_token_index_from_row"] = [ix.get_token_index_from_row(r) for r in rows]
    out["get_doc_index_from_row"] = [ix.get_doc_index_from_row(r) for r in rows]
    dd = [0, 1, 7, docs.shape[0] - 1]
    out["docs"] = dd
    out["get_doc"] = [list(ix.get_doc(d)) for d in dd]
    out["get_doc_length"] = [ix.get_doc_length(d) for d in dd]
    out["get_doc_index"] = [ix.get_doc_index(p) for p in (0, 23, 24, len(ix) - 1)]
    return out


def main():
    ref = load_reference_index()
    docs = make_corpus(**CORPUS)
    seqs = [d.tolist() for d in docs]
    results = {}
    for in_memory in (True, False):                                # index.py:39-66: list path and temp-file (`<l`) path
        ix = ref.FMIndex(); ix.initialize(seqs, in_memory=in_memory)
        results[in_memory] = probe(ix, docs)
    assert results[True] == results[False], "in_memory and file-backed builds disagree"
    with tempfile.TemporaryDirectory() as td:                      # save / load round trip (index.py:186-204)
        ix.save(os.path.join(td, "x"))
        ix2 = ref.FMIndex.load(os.path.join(td, "x"))
        assert probe(ix2, docs) == results[True]
    ora = OracleIndex(seqs, backend="ref")
    got = probe(ora, docs)
    bad = [k for k in results[True] if got[k] != results[True][k]]
    assert not bad, f"OracleIndex differs from seal/index.py in {bad}"
    port = probe(OracleIndex(seqs, backend="port"), docs)
    assert port == results[True], "C-port-backed OracleIndex differs"
    with open(os.path.join(HERE, "index_golden.json"), "w") as f:
        json.dump({"corpus": CORPUS, "probe": results[True]}, f)
    print("OracleIndex == reference seal/index.py on", len(results[True]), "probes; wrote index_golden.json",
          os.path.getsize(os.path.join(HERE, "index_golden.json")), "bytes")


if __name__ == "__main__":
    main()
