"""Evidence aggregation (seal/keys.py:178-497): the oracle restatement and the batched product
implementation against fixtures produced by the REFERENCE FUNCTION ITSELF
(tests/golden/make_keys_golden.py -> keys_golden.json): same document order, same key order, every
float bit-identical."""
import json
import os

import numpy as np
import pytest

from seal_b200.synthetic import make_corpus

GOLD = os.path.join(os.path.dirname(__file__), "golden", "keys_golden.json")


def load_gold():
    with open(GOLD) as f:
        return json.load(f)


def flatten(results, all_ngrams):
    return ([[int(d), v[0], [[list(map(int, k)), s] for k, s in v[1]], [int(t) for t in v[3]],
              [list(map(int, v[4][0])), v[4][1]]] for d, v in results.items()],
            [[list(map(int, k)), s] for k, s in all_ngrams.items()])


def case_ids():
    return [c["name"] for c in load_gold()["cases"]]


@pytest.mark.parametrize("name", case_ids())
def test_oracle_restatement_matches_reference_function(name):
    from oracle.fm_oracle import OracleIndex
    from oracle.keys_oracle import aggregate_evidence_oracle
    g = load_gold()
    docs = make_corpus(**g["corpus"])
    index = OracleIndex([list(map(int, d)) for d in docs])
    c = next(c for c in g["cases"] if c["name"] == name)
    res, alln = aggregate_evidence_oracle([(list(k), s) for k, s in c["keys"]], unigram_scores=c["unigram_scores"],
                                          index=index, **c["kw"])
    got_r, got_a = flatten(res, alln)
    assert got_a == c["all_ngrams"]
    assert got_r == c["results"]


@pytest.fixture(scope="module")
def gpu_index():
    from seal_b200.index import FMIndex
    g = load_gold()
    docs = make_corpus(**g["corpus"])
    idx = FMIndex()
    idx.initialize([list(map(int, d)) for d in docs])
    return idx


@pytest.mark.gpu
@pytest.mark.parametrize("name", case_ids())
def test_product_matches_reference_function(name, gpu_index):
    from seal_b200.keys import aggregate_evidence
    c = next(c for c in load_gold()["cases"] if c["name"] == name)
    res, alln = aggregate_evidence([(list(k), s) for k, s in c["keys"]], unigram_scores=c["unigram_scores"],
                                   index=gpu_index, **c["kw"])
    got_r, got_a = flatten(res, alln)
    assert got_a == c["all_ngrams"]
    assert got_r == c["results"]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(max_occurrences_1=25, sort_by_freq=True), dict(allow_overlaps=True, single_key=0.3)])
def test_product_matches_oracle_on_larger_corpus(kw):
    """4 000 documents, 300 keys (tensor keys as the decode returns them), default 500-document shortlist."""
    import torch
    from oracle.fm_oracle import OracleIndex
    from oracle.keys_oracle import aggregate_evidence_oracle
    from seal_b200.index import FMIndex
    from seal_b200.keys import aggregate_evidence
    docs = make_corpus(n_docs=4000, doc_len=40, n_phrases=3000, seed=9, vocab=3000)
    seqs = [list(map(int, d)) for d in docs]
    ora = OracleIndex(seqs)
    idx = FMIndex(); idx.initialize(seqs)
    rng = np.random.default_rng(3)
    keys = []
    for _ in range(300):
        d = int(rng.integers(0, docs.shape[0])); L = int(rng.integers(1, 9)); a = int(rng.integers(0, docs.shape[1] - L))
        keys.append((torch.tensor(docs[d, a:a + L].astype(np.int64)), float(-rng.exponential(4.0) - 0.01)))
    z = rng.standard_normal(3000) * 3.0
    uni = (z - np.log(np.exp(z).sum())).tolist()
    exp = flatten(*aggregate_evidence_oracle(keys, unigram_scores=uni, index=ora, **kw))
    got = flatten(*aggregate_evidence(keys, unigram_scores=uni, index=idx, **kw))
    assert got[1] == exp[1]
    assert [r[0] for r in got[0]] == [r[0] for r in exp[0]]
    assert got[0] == exp[0]
    assert len(got[0]) > 100


class _BatchedOracleIndex:
    """Test double: the three batched index calls aggregate_evidence makes, answered by the CPU oracle index
    (host-logic check of the product function without a GPU)."""

    def __init__(self, ora):
        self.ora, self.beginnings = ora, ora.beginnings
        self.calls = {"ranges": 0, "locate": 0, "docs": 0}

    def __len__(self):
        return len(self.ora)

    def get_range_batch(self, seqs):
        self.calls["ranges"] += 1
        r = [self.ora.get_range(s) for s in seqs]
        return np.array([a for a, _ in r], dtype=np.uint64), np.array([b for _, b in r], dtype=np.uint64)

    def locate_rows(self, rows):
        self.calls["locate"] += 1
        pos = np.array([self.ora.locate(int(r)) for r in rows], dtype=np.uint64)
        return pos, np.array([self.ora.get_doc_index(int(p)) for p in pos], dtype=np.int64)

    def get_docs(self, ds):
        self.calls["docs"] += 1
        return [self.ora.get_doc(d) for d in ds]


@pytest.mark.parametrize("name", case_ids())
def test_product_host_logic_with_oracle_backed_index(name):
    from oracle.fm_oracle import OracleIndex
    from seal_b200.keys import aggregate_evidence
    g = load_gold()
    docs = make_corpus(**g["corpus"])
    index = _BatchedOracleIndex(OracleIndex([list(map(int, d)) for d in docs]))
    c = next(c for c in g["cases"] if c["name"] == name)
    res, alln = aggregate_evidence([(list(k), s) for k, s in c["keys"]], unigram_scores=c["unigram_scores"],
                                   index=index, **c["kw"])
    got_r, got_a = flatten(res, alln)
    assert got_a == c["all_ngrams"]
    assert got_r == c["results"]
    assert index.calls["locate"] <= 1 and index.calls["docs"] == 1 and index.calls["ranges"] <= 3


# ---- teacher-forced scoring (seal/keys.py:64-176) against the reference functions' own outputs -------
def _tf_gold():
    return load_gold()["teacher_forced"]


def _check_rescore(got, exp, tol):
    assert len(got) == len(exp)
    worst = 0.0
    for a, b in zip(got, exp):
        assert [list(k) for _, k in a] == [k for _, k in b]
        worst = max([worst] + [abs(sa - sb) for (sa, _), (sb, _) in zip(a, b)])
    assert worst < tol, worst
    return worst


def test_keys_oracle_reproduces_reference_rescore_and_unigram_outputs():
    import torch
    from oracle.decode_oracle import make_bart
    from oracle.keys_oracle import rescore_keys_oracle, compute_unigram_scores_oracle
    g = _tf_gold()
    model = make_bart(**g["model"])
    for c in g["rescore"]:
        _check_rescore(rescore_keys_oracle(model, g["inputs"], g["keys"], **c["kw"]), c["out"], 1e-5)
    for c in g["unigram"]:
        lp = compute_unigram_scores_oracle(model, g["inputs"], **c["kw"])
        head = torch.tensor(c["logprobs_head"])
        fin = torch.isfinite(head)
        assert torch.equal(torch.isfinite(lp[:, :64]), fin)
        assert float((lp[:, :64][fin] - head[fin]).abs().max()) < 1e-5
        assert lp.argmax(-1).tolist() == c["row_argmax"]


@pytest.mark.gpu
def test_product_rescore_and_unigram_match_reference_outputs():
    from oracle.decode_oracle import make_bart
    from seal_b200.keys import rescore_keys, compute_unigram_scores
    g = _tf_gold()
    model = make_bart(**g["model"])
    for c in g["rescore"]:
        w = _check_rescore(rescore_keys(model, g["inputs"], g["keys"], **c["kw"]), c["out"], 1e-4)
        print(f"rescore vs reference fixture {c['kw']}: worst |dscore| = {w:.3e}")
    _check_rescore(rescore_keys(model, None, g["keys"][:2]), g["rescore_no_inputs"], 1e-4)
    for c in g["unigram"]:
        lp = compute_unigram_scores(model, g["inputs"], tolist=False, **c["kw"])
        head = np.array(c["logprobs_head"], dtype=np.float64)
        fin = np.isfinite(head)
        assert np.array_equal(np.isfinite(lp[:, :64]), fin)
        assert np.abs(lp[:, :64][fin] - head[fin]).max() < 2e-5
        assert np.abs(lp.max(-1) - np.array(c["row_max"])).max() < 2e-5


def test_product_host_logic_fuzz_against_oracle():
    """Tie-heavy random cases (few distinct scores, repeated keys, every flag): the product function -- batched index
    calls answered by the oracle-backed double, order-defining loops in native code -- must return exactly what the
    oracle restatement returns (which tests/golden/make_keys_golden.py --fuzz checks against the reference function)."""
    from oracle.fm_oracle import OracleIndex
    from oracle.keys_oracle import aggregate_evidence_oracle
    from seal_b200.keys import aggregate_evidence
    rng = np.random.default_rng(77)
    flags = ["sort_by_length", "sort_by_freq", "allow_overlaps", "add_best_unigrams_to_ngrams", "single_key_add_unigrams",
             "unigrams_ignore_free_places"]
    checked = 0
    for case in range(80):
        vocab = int(rng.integers(30, 200))
        docs = make_corpus(n_docs=int(rng.integers(5, 60)), doc_len=int(rng.integers(6, 20)), n_phrases=int(rng.integers(5, 40)),
                           seed=int(rng.integers(0, 1 << 30)), vocab=vocab)
        ora = OracleIndex([list(map(int, d)) for d in docs])
        levels = -np.round(rng.exponential(2.0, size=4), 1) - 0.1
        keys = []
        for _ in range(int(rng.integers(1, 25))):
            d = int(rng.integers(0, docs.shape[0])); L = int(rng.integers(1, 5)); a = int(rng.integers(0, docs.shape[1] - L))
            k = [int(t) for t in docs[d, a:a + L]]
            if rng.random() < 0.15:
                k[-1] = int(rng.integers(4, vocab))
            keys.append((k, float(levels[int(rng.integers(0, len(levels)))])))
        uni = None
        if rng.random() < 0.7:
            z = np.round(rng.standard_normal(vocab), 1)
            uni = (z - np.log(np.exp(z).sum())).tolist()
        kw = {f: bool(rng.random() < 0.4) for f in flags}
        kw.update(use_fm_index_frequency=bool(rng.random() < 0.75), max_occurrences_1=int(rng.choice([1, 3, 50, 1500])),
                  n_docs_complete_score=int(rng.choice([1, 5, 500])), single_key=float(rng.choice([0.0, 0.3, 1.0])),
                  beta=float(rng.choice([0.0, 0.8, 1.0])), alpha=float(rng.choice([1.0, 2.0])),
                  length_penalty=float(rng.choice([0.0, 0.2])), use_top_k_unigrams=int(rng.choice([3, 1000])),
                  max_occurrences_2=int(rng.choice([10, 10_000_000])))
        try:
            exp = flatten(*aggregate_evidence_oracle([(list(k), s) for k, s in keys], unigram_scores=uni, index=ora, **kw))
        except Exception:
            with pytest.raises(Exception):
                aggregate_evidence([(list(k), s) for k, s in keys], unigram_scores=uni, index=_BatchedOracleIndex(ora), **kw)
            continue
        got = flatten(*aggregate_evidence([(list(k), s) for k, s in keys], unigram_scores=uni, index=_BatchedOracleIndex(ora), **kw))
        assert got == exp, (case, kw)
        checked += 1
    assert checked > 50


def test_product_edge_cases_against_oracle():
    """Empty key lists, keys that do not occur, no unigram scores, an empty shortlist, and the IndexError the reference
    raises for use_fm_index_frequency=False without keys: same value or same exception type as the restatement."""
    from oracle.fm_oracle import OracleIndex
    from oracle.keys_oracle import aggregate_evidence_oracle
    from seal_b200.keys import aggregate_evidence
    docs = make_corpus(n_docs=30, doc_len=12, n_phrases=20, seed=3, vocab=80)
    ora = OracleIndex([list(map(int, d)) for d in docs])
    uni = (np.zeros(80) - np.log(80.0)).tolist()
    cases = [([], uni, {}), ([], None, {}), ([([79, 78, 77], -1.0)], uni, {}),
             ([([int(docs[0, 0]), int(docs[0, 1])], -0.5)], None, {}),
             ([([int(docs[0, 0])], -0.5)], uni, dict(n_docs_complete_score=0)),
             ([], uni, dict(use_fm_index_frequency=False))]
    for keys, u, kw in cases:
        try:
            exp, err = flatten(*aggregate_evidence_oracle(list(keys), unigram_scores=u, index=ora, **kw)), None
        except Exception as e:
            exp, err = None, type(e)
        if err is not None:
            with pytest.raises(err):
                aggregate_evidence(list(keys), unigram_scores=u, index=_BatchedOracleIndex(ora), **kw)
        else:
            assert flatten(*aggregate_evidence(list(keys), unigram_scores=u, index=_BatchedOracleIndex(ora), **kw)) == exp
