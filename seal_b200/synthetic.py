"""Seeded synthetic inputs for the benchmark and the parity tests (SURVEY.md §8d recipe).

Corpus: a phrase-dictionary generator over BART's id space — collocations repeat verbatim and
phrase boundaries have huge fan-out, the two properties of real text that decide SA-range widths
and the skew of the per-row successor sets.  Pure numpy; no reference code involved.
"""
import numpy as np

VOCAB = 50265          # len(BART tokenizer), seal/retrieval.py:570
EOS = 2
PAD = 1
BOS = 0


def make_corpus(n_docs=100_000, doc_len=100, n_phrases=200_000, seed=1234, vocab=VOCAB):
    """Returns int32 [n_docs, doc_len]; each doc = doc_len-1 body tokens + </s> (build_fm_index.py:132)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_tok = vocab - 4                                   # usable ids 4..vocab-1
    perm = rng.permutation(n_tok) + 4                   # rank -> token id
    p = 1.0 / (np.arange(n_tok) + 1.0)
    cdf = np.cumsum(p / p.sum())
    plen = rng.integers(2, 13, size=n_phrases)
    ptok = perm[np.minimum(np.searchsorted(cdf, rng.random(int(plen.sum()))), n_tok - 1)]
    pstart = np.zeros(n_phrases + 1, dtype=np.int64)
    np.cumsum(plen, out=pstart[1:])
    q = 1.0 / (np.arange(n_phrases) + 1.0)
    pcdf = np.cumsum(q / q.sum())
    per_doc = max(1, (doc_len + 1) // 2)                # phrases have >= 2 tokens
    ph = np.minimum(np.searchsorted(pcdf, rng.random((n_docs, per_doc))), n_phrases - 1)
    body = doc_len - 1
    docs = np.empty((n_docs, doc_len), dtype=np.int32)
    # vectorised concatenation: gather each doc's phrases then cut to `body` tokens
    lens = plen[ph]                                     # [n_docs, per_doc]
    ends = np.cumsum(lens, axis=1)
    starts = ends - lens
    pos = np.arange(body)[None, :]                      # [1, body]
    # phrase slot of every body position
    slot = (pos[:, :, None] >= ends[:, None, :]).sum(axis=2) if n_docs * body * per_doc <= 5e7 else None
    if slot is None:
        slot = np.empty((n_docs, body), dtype=np.int64)
        chunk = max(1, int(5e7 // (body * per_doc)))
        for a in range(0, n_docs, chunk):
            b = min(n_docs, a + chunk)
            slot[a:b] = (pos[:, :, None] >= ends[a:b, None, :]).sum(axis=2)
    slot = np.minimum(slot, per_doc - 1)
    rows = np.arange(n_docs)[:, None]
    within = pos - starts[rows, slot]
    docs[:, :body] = ptok[pstart[ph[rows, slot]] + within]
    docs[:, body] = EOS
    return docs


def corpus_symbols(docs, shift=10):
    """seal/index.py:50-53: per-document reversal, +SHIFT, concatenation -> u64 symbol stream."""
    return (docs[:, ::-1].astype(np.uint64) + np.uint64(shift)).reshape(-1)


def make_queries(n_queries=1000, min_len=12, max_len=28, seed=4321, vocab=VOCAB):
    """Encoder inputs: <s> + random ids + </s>, right-padded with <pad>; returns (ids, mask) int64."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(min_len, max_len + 1, size=n_queries)
    S = int(lens.max())
    ids = np.full((n_queries, S), PAD, dtype=np.int64)
    mask = np.zeros((n_queries, S), dtype=np.int64)
    for i, l in enumerate(lens):
        ids[i, 0] = BOS
        ids[i, 1:l - 1] = rng.integers(4, vocab, size=l - 2)
        ids[i, l - 1] = EOS
        mask[i, :l] = 1
    return ids, mask
