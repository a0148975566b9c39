"""CPU: the plain-C oracle (oracle/fm_oracle.c) against the committed golden vectors that the
compiled reference produced (tests/golden/make_golden.py), and against the compiled reference
itself where oracle/_ref exists."""
import os

import numpy as np
import pytest

from oracle.fm_oracle import PortFM, RefFM, ref_available, OracleIndex

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fm_golden.npz"))
CASES = ["keeper", "toy", "rand5k", "phrase"]


def check_against_golden(fm, name):
    size = int(G[f"{name}.size"])
    assert fm.size() == size
    fl, fh = G[f"{name}.first_lo"], G[f"{name}.first_hi"]
    syms = np.arange(len(fl), dtype=np.uint64)
    ol, oh = fm.backward_search_step_batch(syms, np.zeros_like(syms), np.full_like(syms, size))
    assert np.array_equal(ol, fl) and np.array_equal(oh, fh)
    wsym, wlo, whi = G[f"{name}.walk_sym"], G[f"{name}.walk_lo"], G[f"{name}.walk_hi"]
    dc_off, dc = G[f"{name}.dc_off"], G[f"{name}.dc"]
    W, D = wsym.shape
    k = 0
    for w in range(W):
        lo, hi = 0, size
        for d in range(D):
            lo, hi = fm.backward_search_step(int(wsym[w, d]), lo, hi)
            assert (lo, hi) == (int(wlo[w, d]), int(whi[w, d]))
            got = fm.distinct_count(lo, hi + 1) if hi + 1 >= lo else np.zeros(0, dtype=np.uint64)
            assert np.array_equal(got, dc[int(dc_off[k]):int(dc_off[k + 1])])
            k += 1
    for row, exp in zip(G[f"{name}.loc_rows"], G[f"{name}.loc"]):
        assert fm.locate(int(row)) == int(exp)
    eo, ex = G[f"{name}.ext_off"], G[f"{name}.ext"]
    for i, (b, e) in enumerate(zip(G[f"{name}.ext_b"], G[f"{name}.ext_e"])):
        assert np.array_equal(fm.extract_text(int(b), int(e)), ex[int(eo[i]):int(eo[i + 1])])


@pytest.mark.parametrize("name", CASES)
def test_port_matches_reference_goldens(name):
    check_against_golden(PortFM(G[f"{name}.text"]), name)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", CASES)
def test_compiled_reference_reproduces_goldens(name):
    check_against_golden(RefFM(G[f"{name}.text"]), name)


def test_first_step_quirk_is_pinned():
    """get_range's first step passes r = size(); at least one golden case must
    show hi one too large relative to the in-contract call r = size()-1."""
    seen = 0
    for name in CASES:
        fm = PortFM(G[f"{name}.text"])
        size = fm.size()
        fl, fh = G[f"{name}.first_lo"], G[f"{name}.first_hi"]
        for s in range(len(fl)):
            lo, hi = fm.backward_search_step(s, 0, size - 1)
            if (lo, hi) != (int(fl[s]), int(fh[s])):
                assert lo == int(fl[s]) and hi + 1 == int(fh[s])
                seen += 1
    assert seen >= 1


def test_oracle_index_api_matches_seal_semantics():
    docs = G["phrase.docs"]
    idx = OracleIndex([d.tolist() for d in docs], backend="port")
    assert len(idx) == docs.size and idx.n_docs == len(docs)
    # every document's trigram must be found; its continuation set contains the true next token
    for d in docs[:20]:
        d = d.tolist()
        lo, hi = idx.get_range(d[3:6])
        assert hi - lo >= 1
        assert d[6] in idx.get_continuations(d[3:6])
    assert idx.get_range([]) == (0, docs.size + 2)
    # get_doc returns the document as stored (reversed text walked backwards = forward order)
    assert idx.get_doc(3) == docs[3].tolist()
    # locate: position of the n-gram's last token in reversed-text coordinates -> same document
    lo, hi = idx.get_range(docs[7].tolist()[2:6])
    assert 7 in {idx.get_doc_index_from_row(r) for r in range(lo, hi)}


def _decode_gold():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "decode_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", range(9))
def test_decode_oracle_reproduces_reference_decode_outputs(case):
    """decode_golden.json = outputs of the reference's OWN seal/beam_search.py (run unmodified by
    tests/golden/make_decode_golden.py): same hypotheses, same order; scores to 1e-5 (the fixture was
    produced by eager fp32 PyTorch on another CPU)."""
    import torch
    from oracle.decode_oracle import make_bart, fm_index_generate_oracle
    from seal_b200.synthetic import make_corpus
    g = _decode_gold()
    c = g["cases"][case]
    docs = make_corpus(**g["corpus"])
    ora = OracleIndex([d.tolist() for d in docs])
    model = make_bart(**g["model"])
    out = fm_index_generate_oracle(model, ora, torch.tensor(c["input_ids"]), torch.tensor(c["attention_mask"]), **c["kw"])
    assert len(out) == len(c["hyps"])
    for got, exp in zip(out, c["hyps"]):
        assert [list(t) for _, t, _ in got] == [t for _, t in exp]
        assert max((abs(s - e[0]) for (s, _, _), e in zip(got, exp)), default=0.0) < 1e-5


def test_cached_stepper_equals_reforwarding_stepper():
    """The KV-cached driver of HF BART (what the reference runs: use_cache=True + _reorder_cache, seal/beam_search.py:
    331-332,483) returns the same hypotheses as the prefix re-forwarding one the fixtures were pinned with."""
    import torch
    from oracle.decode_oracle import make_bart, fm_index_generate_oracle
    from oracle.fm_oracle import OracleIndex
    from seal_b200.synthetic import make_corpus
    docs = make_corpus(n_docs=300, doc_len=30, n_phrases=600, seed=3, vocab=2000)
    ora = OracleIndex([d.tolist() for d in docs])
    model = make_bart(seed=0, layers=2, vocab=2000, d_model=128)
    rng = np.random.default_rng(12)
    ids = torch.tensor(rng.integers(4, 2000, size=(3, 10)), dtype=torch.long); ids[:, 0] = 0; ids[:, -1] = 2
    am = torch.ones_like(ids); ids[1, 7:] = 1; ids[1, 6] = 2; am[1, 7:] = 0
    kw = dict(num_beams=5, min_length=7, max_length=7, length_penalty=0.0)
    a = fm_index_generate_oracle(model, ora, ids, am, **kw)
    b = fm_index_generate_oracle(model, ora, ids, am, use_cache=True, **kw)
    for qa, qb in zip(a, b):
        assert [tuple(t) for _, t, _ in qa] == [tuple(t) for _, t, _ in qb]
        assert all(abs(x[0] - y[0]) < 1e-5 for x, y in zip(qa, qb))
