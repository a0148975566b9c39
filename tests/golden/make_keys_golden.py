"""Generates tests/golden/keys_golden.json by running the REFERENCE's own aggregate_evidence
(/root/reference/seal/keys.py:178-497, imported unmodified; its `seal` and `more_itertools` imports are
satisfied by stub modules -- the function only needs an object with the seal.index.FMIndex methods,
here oracle.fm_oracle.OracleIndex on the compiled reference FM-index).  Run in the build container only:

    python tests/golden/make_keys_golden.py

Each case stores the corpus recipe, the inputs and the complete return value (document order, key
order, every float) so that tests can rebuild the index and compare exactly."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.fm_oracle import OracleIndex  # noqa: E402
from seal_b200.synthetic import make_corpus  # noqa: E402

CORPUS = dict(n_docs=300, doc_len=30, n_phrases=400, seed=77, vocab=600)


def load_reference_keys():
    stub = types.ModuleType("seal"); stub.FMIndex = OracleIndex
    mi = types.ModuleType("more_itertools"); mi.chunked = lambda it, n: (it[i:i + n] for i in range(0, len(it), n))
    sys.modules.setdefault("seal", stub); sys.modules.setdefault("more_itertools", mi)
    spec = importlib.util.spec_from_file_location("ref_keys", "/root/reference/seal/keys.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


def make_inputs(docs, seed, n_keys=60, vocab=600, with_unigrams=True):
    """Keys as the decode would deliver them: n-grams cut out of documents (so they occur), a few that do
    not occur, duplicates with different scores, single tokens; scores = negative log-probs."""
    rng = np.random.default_rng(seed)
    keys = []
    for _ in range(n_keys):
        d = int(rng.integers(0, docs.shape[0])); L = int(rng.integers(1, 7)); a = int(rng.integers(0, docs.shape[1] - L))
        k = [int(t) for t in docs[d, a:a + L]]
        if rng.random() < 0.1:
            k[-1] = int(rng.integers(4, vocab))                 # probably absent from the corpus
        keys.append((k, float(-rng.exponential(3.0) - 0.05)))
    keys += [(list(keys[i][0]), float(keys[i][1] - 0.5)) for i in range(0, 10, 3)]     # repeated keys
    uni = None
    if with_unigrams:
        z = rng.standard_normal(vocab) * 2.0
        uni = (z - np.log(np.exp(z).sum())).tolist()
    return keys, uni


CASES = [
    dict(name="default", seed=1, kw={}),
    dict(name="no_unigrams", seed=2, kw={}, with_unigrams=False),
    dict(name="rare_freq_split", seed=3, kw=dict(max_occurrences_1=3, n_docs_complete_score=25)),
    dict(name="sort_by_length", seed=4, kw=dict(sort_by_length=True, max_occurrences_1=40)),
    dict(name="sort_by_freq", seed=5, kw=dict(sort_by_freq=True, max_occurrences_1=40)),
    dict(name="overlaps_single_key", seed=6, kw=dict(allow_overlaps=True, single_key=0.5, single_key_add_unigrams=True)),
    dict(name="no_fm_frequency", seed=7, kw=dict(use_fm_index_frequency=False, alpha=1.5)),
    dict(name="add_best_unigrams", seed=8, kw=dict(add_best_unigrams_to_ngrams=True, use_top_k_unigrams=50,
                                                   length_penalty=0.1, beta=0.5, smoothing=2.0)),
    dict(name="ignore_free_places", seed=9, kw=dict(unigrams_ignore_free_places=True, max_occurrences_2=200,
                                                    n_docs_complete_score=10)),
]


def dump_result(results, all_ngrams):
    return {"results": [[int(d), v[0], [[list(map(int, k)), s] for k, s in v[1]], [int(t) for t in v[3]],
                         [list(map(int, v[4][0])), v[4][1]]] for d, v in results.items()],
            "all_ngrams": [[list(map(int, k)), s] for k, s in all_ngrams.items()]}


class Model413Proxy:
    """What seal/keys.py:64-176 needs from a transformers-4.13 BART: parameters(), config, forward, and the
    private `_prepare_encoder_decoder_kwargs_for_generation(input_ids, kwargs)` (4.13 signature)."""
    def __init__(self, model): self.m, self.config = model, model.config
    def parameters(self): return self.m.parameters()
    def __call__(self, **kw): return self.m(**kw)
    def _prepare_encoder_decoder_kwargs_for_generation(self, input_ids, kw):
        kw["encoder_outputs"] = self.m.get_encoder()(input_ids=input_ids, attention_mask=kw["attention_mask"], return_dict=True)
        return kw


def teacher_forced_cases(ref):
    """rescore_keys / compute_unigram_scores of the reference (keys.py:64-176) on the tiny seeded BART the
    GPU tests use; checked against oracle/keys_oracle.py and stored."""
    import torch
    from oracle.decode_oracle import make_bart
    from oracle.keys_oracle import rescore_keys_oracle, compute_unigram_scores_oracle
    MODEL = dict(seed=0, layers=2, vocab=2000, d_model=128)
    model = make_bart(**MODEL)
    proxy = Model413Proxy(model)
    rng = np.random.default_rng(12)
    inputs = [[0] + rng.integers(4, 2000, size=int(rng.integers(3, 9))).tolist() + [2] for _ in range(5)]
    keys = []
    for q in range(5):
        n = int(rng.integers(0, 7)) if q != 2 else 0
        kk = []
        for _ in range(n):
            toks = rng.integers(4, 2000, size=int(rng.integers(1, 8))).tolist()
            u = rng.random()
            if u < 0.3: toks = [0] + toks
            if u > 0.6: toks = toks + [2]
            kk.append((float(rng.random()), toks) if rng.random() < 0.5 else toks)
        keys.append(kk)
    out = {"model": MODEL, "inputs": inputs, "keys": keys, "rescore": [], "unigram": []}
    for kw in [dict(), dict(length_penalty=1.0), dict(prefix=[7, 9]), dict(strip_from_bos=[0], strip_from_eos=[2]), dict(batch_size=3)]:
        got = ref.rescore_keys(proxy, [list(i) for i in inputs], [list(k) for k in keys], **kw)
        exp = rescore_keys_oracle(model, inputs, keys, **kw)
        worst = 0.0
        for a, b in zip(got, exp):
            assert [k for _, k in a] == [k for _, k in b]
            worst = max([worst] + [abs(sa - sb) for (sa, _), (sb, _) in zip(a, b)])
        print("rescore", kw, "oracle vs reference worst |dscore|", worst)
        assert worst < 2e-5
        out["rescore"].append({"kw": kw, "out": [[[s, k] for s, k in q] for q in got]})
    got0 = ref.rescore_keys(proxy, None, [list(k) for k in keys[:2]])
    out["rescore_no_inputs"] = [[[s, k] for s, k in q] for q in got0]
    for kw in [dict(), dict(temperature=0.7), dict(prefix=[11])]:
        got = ref.compute_unigram_scores(proxy, [list(i) for i in inputs], None, tolist=False, **kw)
        exp = compute_unigram_scores_oracle(model, inputs, **kw)
        fin = torch.isfinite(exp)
        assert torch.equal(torch.isfinite(got), fin)
        err = float((got[fin] - exp[fin]).abs().max())
        print("unigram", kw, "oracle vs reference max |dlogprob|", err)
        assert err < 1e-5
        out["unigram"].append({"kw": kw, "logprobs_head": got[:, :64].tolist(), "row_max": got.max(-1).values.tolist(),
                               "row_argmax": got.argmax(-1).tolist()})
    return out


def main():
    ref = load_reference_keys()
    docs = make_corpus(**CORPUS)
    index = OracleIndex([list(map(int, d)) for d in docs], backend="ref")
    out = {"corpus": CORPUS, "cases": []}
    for c in CASES:
        keys, uni = make_inputs(docs, c["seed"], with_unigrams=c.get("with_unigrams", True))
        kw = dict(n_docs_complete_score=40); kw.update(c["kw"])          # keep the fixture small
        results, all_ngrams = ref.aggregate_evidence([(list(k), s) for k, s in keys], unigram_scores=uni, index=index, **kw)
        rec = {"name": c["name"], "kw": kw, "keys": keys, "unigram_scores": uni}
        rec.update(dump_result(results, all_ngrams))
        print(c["name"], "docs scored:", len(rec["results"]), "keys kept:", len(rec["all_ngrams"]),
              "top:", rec["results"][0][:2] if rec["results"] else None)
        out["cases"].append(rec)
    out["teacher_forced"] = teacher_forced_cases(ref)
    with open(os.path.join(HERE, "keys_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote keys_golden.json", os.path.getsize(os.path.join(HERE, "keys_golden.json")), "bytes")


def fuzz(n_cases):
    """Randomised cross-check (nothing stored): the reference function vs the oracle restatement on small corpora with
    many score ties, repeated keys and every flag combination -- the orders the fixtures cannot pin exhaustively."""
    import itertools
    from oracle.keys_oracle import aggregate_evidence_oracle
    ref = load_reference_keys()
    rng = np.random.default_rng(2024)
    flags = ["sort_by_length", "sort_by_freq", "allow_overlaps", "use_fm_index_frequency", "add_best_unigrams_to_ngrams",
             "single_key_add_unigrams", "unigrams_ignore_free_places"]
    bad = 0
    for case in range(n_cases):
        vocab = int(rng.integers(30, 200))
        docs = make_corpus(n_docs=int(rng.integers(5, 60)), doc_len=int(rng.integers(6, 20)), n_phrases=int(rng.integers(5, 40)),
                           seed=int(rng.integers(0, 1 << 30)), vocab=vocab)
        index = OracleIndex([list(map(int, d)) for d in docs], backend="ref")
        levels = -np.round(rng.exponential(2.0, size=4), 1) - 0.1               # few distinct scores -> ties everywhere
        keys = []
        for _ in range(int(rng.integers(1, 25))):
            d = int(rng.integers(0, docs.shape[0])); L = int(rng.integers(1, 5)); a = int(rng.integers(0, docs.shape[1] - L))
            k = [int(t) for t in docs[d, a:a + L]]
            if rng.random() < 0.15:
                k[-1] = int(rng.integers(4, vocab))
            keys.append((k, float(levels[int(rng.integers(0, len(levels)))])))
        uni = None
        if rng.random() < 0.7:
            z = np.round(rng.standard_normal(vocab), 1)                          # ties among unigram scores too
            uni = (z - np.log(np.exp(z).sum())).tolist()
        kw = {f: bool(rng.random() < 0.4) for f in flags}
        kw["use_fm_index_frequency"] = not kw["use_fm_index_frequency"] if rng.random() < 0.5 else True
        kw.update(max_occurrences_1=int(rng.choice([1, 3, 50, 1500])), n_docs_complete_score=int(rng.choice([1, 5, 500])),
                  single_key=float(rng.choice([0.0, 0.3, 1.0])), beta=float(rng.choice([0.0, 0.8, 1.0])),
                  alpha=float(rng.choice([1.0, 2.0])), length_penalty=float(rng.choice([0.0, 0.2])),
                  use_top_k_unigrams=int(rng.choice([3, 1000])), max_occurrences_2=int(rng.choice([10, 10_000_000])))
        try:
            exp = dump_result(*ref.aggregate_evidence([(list(k), s) for k, s in keys], unigram_scores=uni, index=index, **kw))
        except Exception as e:                                                   # the oracle must fail the same way
            try:
                aggregate_evidence_oracle([(list(k), s) for k, s in keys], unigram_scores=uni, index=index, **kw)
                print("case", case, "reference raised", type(e).__name__, "but the oracle did not"); bad += 1
            except Exception as e2:
                if type(e2) is not type(e):
                    print("case", case, "different exceptions", type(e).__name__, type(e2).__name__); bad += 1
            continue
        got = dump_result(*aggregate_evidence_oracle([(list(k), s) for k, s in keys], unigram_scores=uni, index=index, **kw))
        if got != exp:
            bad += 1
            print("case", case, "MISMATCH", kw)
    print(f"fuzz: {n_cases} cases, {bad} mismatches")
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--fuzz":
        sys.exit(1 if fuzz(int(sys.argv[2])) else 0)
    main()
