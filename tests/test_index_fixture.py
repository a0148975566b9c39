"""seal/index.py API parity against answers produced by the REFERENCE'S OWN seal/index.py + fm_index.cpp
(tests/golden/make_index_golden.py -> index_golden.json): the oracle index on CPU, the product index on GPU."""
import json
import os
import sys

import pytest

from seal_b200.synthetic import make_corpus

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
GOLD = os.path.join(os.path.dirname(__file__), "golden", "index_golden.json")


def replay(ix, g):
    """Re-asks every stored question (arguments come from the fixture, not from a generator)."""
    p = g["probe"]
    out = {"len": len(ix), "n_docs": ix.n_docs, "size": ix.size(), "beginnings_tail": list(ix.beginnings[-3:]),
           "occurring_distinct": list(ix.occurring_distinct), "occurring_counts": list(ix.occurring_counts)}
    seqs = p["seqs"]
    out["seqs"] = seqs
    out["get_range"] = [list(ix.get_range(s)) for s in seqs]
    out["get_count"] = [ix.get_count(s) for s in seqs]
    out["get_continuations"] = [list(ix.get_continuations(s)) for s in seqs[:30]]
    out["get_doc_indices"] = [list(ix.get_doc_indices(s))[:50] for s in seqs[1:30]]
    rr = p["ranges"]
    out["ranges"] = rr
    out["get_distinct"] = [list(ix.get_distinct(a, b)) for a, b in rr]
    out["get_distinct_count"] = [[list(x) for x in ix.get_distinct_count(a, b)] for a, b in rr]
    out["get_distinct_count_multi"] = [[list(x) for x in q] for q in ix.get_distinct_count_multi([a for a, _ in rr], [b for _, b in rr])]
    out["rows"] = p["rows"]
    out["get_token_index_from_row"] = [ix.get_token_index_from_row(r) for r in p["rows"]]
    out["get_doc_index_from_row"] = [ix.get_doc_index_from_row(r) for r in p["rows"]]
    out["docs"] = p["docs"]
    out["get_doc"] = [list(ix.get_doc(d)) for d in p["docs"]]
    out["get_doc_length"] = [ix.get_doc_length(d) for d in p["docs"]]
    out["get_doc_index"] = [ix.get_doc_index(q) for q in (0, 23, 24, len(ix) - 1)]
    return out


def load():
    with open(GOLD) as f:
        g = json.load(f)
    docs = make_corpus(**g["corpus"])
    return g, [d.tolist() for d in docs]


def diff(got, exp):
    return [k for k in exp if json.loads(json.dumps(got[k])) != exp[k]]


@pytest.mark.parametrize("backend", ["port", "auto"])
def test_oracle_index_answers_like_reference_index_py(backend):
    from oracle.fm_oracle import OracleIndex
    g, seqs = load()
    assert diff(replay(OracleIndex(seqs, backend=backend), g), g["probe"]) == []


@pytest.mark.gpu
@pytest.mark.parametrize("in_memory", [True, False])
def test_product_index_answers_like_reference_index_py(in_memory, tmp_path):
    from seal_b200.index import FMIndex
    g, seqs = load()
    ix = FMIndex(); ix.initialize(seqs, in_memory=in_memory)
    assert diff(replay(ix, g), g["probe"]) == []
    ix.save(str(tmp_path / "x"))                                  # index.py:186-204 round trip
    assert diff(replay(FMIndex.load(str(tmp_path / "x")), g), g["probe"]) == []
