"""GPU parity of the teacher-forced scoring path (seal/keys.py rescore_keys,
compute_unigram_scores) against the torch restatement on transformers' BART in eager fp32."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def setup(vocab=2000, layers=2, d_model=128):
    import torch
    from oracle.decode_oracle import make_bart
    model = make_bart(seed=0, layers=layers, vocab=vocab, d_model=d_model)
    rng = np.random.default_rng(12)
    inputs = [[0] + rng.integers(4, vocab, size=int(rng.integers(3, 9))).tolist() + [2] for _ in range(5)]
    keys = []
    for q in range(5):
        n = int(rng.integers(0, 7)) if q != 2 else 0             # one query without keys
        kk = []
        for _ in range(n):
            toks = rng.integers(4, vocab, size=int(rng.integers(1, 8))).tolist()
            u = rng.random()
            if u < 0.3: toks = [0] + toks                         # leading bos (stripped / masked)
            if u > 0.6: toks = toks + [2]                         # trailing eos
            kk.append((float(rng.random()), toks) if rng.random() < 0.5 else toks)
        keys.append(kk)
    return model, inputs, keys


@pytest.mark.parametrize("kw", [dict(), dict(length_penalty=1.0), dict(prefix=[7, 9]), dict(strip_from_bos=[0], strip_from_eos=[2])])
def test_rescore_keys_vs_oracle(kw):
    from oracle.keys_oracle import rescore_keys_oracle
    from seal_b200.keys import rescore_keys
    model, inputs, keys = setup()
    exp = rescore_keys_oracle(model, inputs, keys, **kw)
    got = rescore_keys(model, inputs, keys, **kw)
    assert len(got) == len(exp)
    worst = 0.0
    for a, b in zip(got, exp):
        assert [k for _, k in a] == [k for _, k in b]
        for (sa, _), (sb, _) in zip(a, b):
            worst = max(worst, abs(sa - sb))
    print(f"rescore {kw}: worst |dscore| = {worst:.3e}")
    assert worst < 1e-4
    # inputs=None path (keys.py:70-73)
    exp0 = rescore_keys_oracle(model, None, keys[:2])
    got0 = rescore_keys(model, None, keys[:2])
    for a, b in zip(got0, exp0):
        for (sa, ka), (sb, kb) in zip(a, b):
            assert ka == kb and abs(sa - sb) < 1e-4


@pytest.mark.parametrize("kw", [dict(), dict(temperature=0.7), dict(prefix=[11])])
def test_unigram_scores_vs_oracle(kw):
    from oracle.keys_oracle import compute_unigram_scores_oracle
    from seal_b200.keys import compute_unigram_scores
    model, inputs, _ = setup()
    exp = compute_unigram_scores_oracle(model, inputs, **kw).numpy()
    got = compute_unigram_scores(model, inputs, tolist=False, **kw)
    fin = np.isfinite(exp)
    assert np.array_equal(np.isfinite(got), fin)
    err = np.abs(got[fin] - exp[fin]).max()
    print(f"unigram {kw}: max |dlogprob| = {err:.3e}")
    assert err < 2e-5
    assert isinstance(compute_unigram_scores(model, inputs[:1])[0], list)


def test_rescore_keys_bart_large():
    from oracle.decode_oracle import make_bart
    from oracle.keys_oracle import rescore_keys_oracle
    from seal_b200.keys import rescore_keys
    model = make_bart(seed=0)
    rng = np.random.default_rng(3)
    inputs = [[0] + rng.integers(4, 50265, size=6).tolist() + [2] for _ in range(2)]
    keys = [[rng.integers(4, 50265, size=int(rng.integers(2, 9))).tolist() for _ in range(4)] for _ in range(2)]
    exp = rescore_keys_oracle(model, inputs, keys)
    got = rescore_keys(model, inputs, keys)
    worst = max(abs(sa - sb) for a, b in zip(got, exp) for (sa, _), (sb, _) in zip(a, b))
    print(f"bart-large rescore: worst |dscore| = {worst:.3e}")
    assert worst < 1e-4
