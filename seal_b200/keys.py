"""Drop-in for ``seal.keys`` (/root/reference/seal/keys.py): the decoder-side helpers that run right
after every generate pass -- ``rescore_keys`` (:64-141) and
``compute_unigram_scores`` (:145-176), on the kernels behind ``sealdec_teacher_forced``
(include/sealdec.h) -- and the evidence aggregation (§8f rank 2) ``aggregate_evidence`` (:178-497),
whose FM-index accesses are three batched GPU launches.  Same signatures and return values.
No CPU path."""
import ctypes as C
from typing import List, Optional

import sys

import numpy as np

from ._lib import lib, check
from .beam_search import _engine_for


def strip(seq, symbols_start, symbols_end):                     # keys.py:53-61
    i = 0
    while i < len(seq) and seq[i] in symbols_start:
        i += 1
    j = len(seq)
    while j > i and seq[j - 1] in symbols_end:
        j -= 1
    return seq[i:j]


def _pad_inputs(batch_in, pad):
    maxlen = max(len(i) for i in batch_in)
    ids = np.full((len(batch_in), maxlen), pad, dtype=np.int64)
    for r, i in enumerate(batch_in):
        ids[r, :len(i)] = i
    return ids, (ids != pad).astype(np.int64)                   # keys.py:78-82


def _teacher_forced(eng, ids, mask, dec, row_query, temperature=1.0, full_pos=-1):
    N, T = dec.shape
    Q, S = ids.shape
    V = int(eng.config.vocab_size)
    out = np.zeros((N, max(T - 1, 1)), dtype=np.float32)
    full = np.empty((N, V), dtype=np.float32) if full_pos >= 0 else None
    rq = np.ascontiguousarray(row_query, dtype=np.int32)
    check(lib.sealdec_teacher_forced(eng._h, ids.ctypes.data, mask.ctypes.data, Q, S, dec.ctypes.data, rq.ctypes.data, N, T,
                                     C.c_float(temperature), out.ctypes.data if T > 1 else None, full_pos,
                                     full.ctypes.data if full is not None else None))
    return out[:, :T - 1], full


def rescore_keys(model, inputs, list_of_decoded, batch_size=100, length_penalty=0.0, progress_bar=False, prefix=[],
                 strip_from_bos=[], strip_from_eos=[]):
    """keys.py:64-141.  `batch_size` is accepted for signature compatibility; the native pass chunks rows itself."""
    eng = _engine_for(model)
    cfg = eng.config
    if inputs is None:                                                           # :70-73
        batch_in = [[cfg.bos_token_id, cfg.eos_token_id]] * len(list_of_decoded)
    else:
        batch_in = [list(i) for i in inputs]
    list_of_decoded = [[x[1] if isinstance(x[0], float) else x for x in xx] for xx in list_of_decoded]   # :75
    ids, mask = _pad_inputs(batch_in, cfg.pad_token_id)
    rows, row_query, orig = [], [], []
    for idx, ddi in enumerate(list_of_decoded):                                  # :87-104
        for di in ddi:
            di = list(di.tolist() if hasattr(di, "tolist") else di)
            stripped = [cfg.decoder_start_token_id] + list(prefix) + strip(di, strip_from_bos, strip_from_eos)
            rows.append(stripped); row_query.append(idx); orig.append(di)
    all_out = {i: [] for i in range(len(list_of_decoded))}
    if not rows:
        return [v for k, v in sorted(all_out.items())]
    T = max(len(r) for r in rows)
    dec = np.full((len(rows), T), cfg.pad_token_id, dtype=np.int64)              # :110-116
    for r, toks in enumerate(rows):
        dec[r, :len(toks)] = toks
    lp, _ = _teacher_forced(eng, ids, mask, dec, row_query)
    lp[dec[:, 1:] < 2] = 0.0                                                     # :132
    lp = lp[:, len(prefix):].sum(-1, dtype=np.float32)                           # :133-134 (the reference sums the fp32 tensor)
    for q, di, ll in zip(row_query, orig, lp.tolist()):
        all_out[q].append((ll / (len(di) ** length_penalty), di))                # :138-139
    return [v for k, v in sorted(all_out.items())]


def compute_unigram_scores(model, inputs, index=None, tokenizer=None, tolist=True, temperature=1.0, prefix=[]):
    """keys.py:145-176: full-vocabulary log-probs of the first decoded position (after `prefix`)."""
    eng = _engine_for(model)
    cfg = eng.config
    if isinstance(inputs[0], str):
        batch = tokenizer(inputs, padding=True, return_tensors="np")
        ids = np.ascontiguousarray(batch["input_ids"], dtype=np.int64)
        mask = np.ascontiguousarray(batch["attention_mask"], dtype=np.int64)
    else:
        ids, mask = _pad_inputs([list(i) for i in inputs], cfg.pad_token_id)
    dec = np.full((ids.shape[0], 1 + len(prefix)), cfg.decoder_start_token_id, dtype=np.int64)   # :164-166
    for i, t in enumerate(prefix, start=1):
        dec[:, i] = t
    _, full = _teacher_forced(eng, ids, mask, dec, np.arange(ids.shape[0]), temperature=temperature, full_pos=len(prefix))
    return full.tolist() if tolist else full


# ------------------------------------------------------------------------------------------------
# Evidence aggregation (/root/reference/seal/keys.py:178-497).
#
# The reference walks the FM-index one call at a time from Python: get_count per key (twice), per
# vocabulary entry, then locate + get_doc_index per SA row (up to max_occurrences_1 rows per key) and
# get_doc per shortlisted document.  Here every index access of a call is one of three batched GPU
# launches -- all interval searches (sealfm_backward_search_multi), all row locations (sealfm_locate)
# and all document extractions (sealfm_extract_text); the ordering / set logic that defines the
# result stays on the host, arranged around those three batches.  Same signature, same return value
# (document order, key order and every float identical to the reference's).
# ------------------------------------------------------------------------------------------------
import math as _math
from collections import Counter as _Counter


class _Evidence:
    def __init__(self, index, p):
        self.index, self.p = index, p
        self.ntokens = float(index.beginnings[-1])                                          # keys.py:193
        self.range_of = {}                       # token tuple -> (lo, hi) half-open SA range

    # -- batch 1: SA ranges ------------------------------------------------------------------------
    def need_ranges(self, seqs):
        todo = [s for s in dict.fromkeys(tuple(s) for s in seqs) if s not in self.range_of]
        if todo:
            lo, hi = self.index.get_range_batch([list(s) for s in todo])
            for s, l, h in zip(todo, lo.tolist(), hi.tolist()):
                self.range_of[s] = (l, h)

    def count(self, seq):
        lo, hi = self.range_of[tuple(seq)]
        return hi - lo

    # -- scalar scoring (kept in Python floats: the reference's exact arithmetic) ------------------
    def contrast(self, sr, count):                                                          # :220-223, :250-253
        p = self.p
        snr = _math.log((count + p["smoothing"]) / (self.ntokens + p["smoothing"]))
        return (sr + _math.log(1 - _math.exp(snr))) - (snr + _math.log(1 - _math.exp(sr)))

    def damp(self, types, score, seen):                                                     # :186-191
        if not seen:
            return score
        types = set(types)
        beta = self.p["beta"]
        return (1.0 - beta + (beta * len(types.difference(seen)) / len(types))) * score

    def key_score(self, key, sr, cutoff):                                                   # :208-235
        p = self.p
        c = self.count(key)
        if c == 0:
            return 0.0
        decay = (1.0 - p["length_penalty"]) ** (len(key) - 1.0)
        if p["use_fm_index_frequency"]:
            sc = max(self.contrast((sr - 1e-10) * decay, c), 0.0)
        else:
            sc = max(sr - cutoff, 0.0) * decay
        return sc ** p["alpha"]

    def unigram_table(self, unigram_scores, given, cutoff):                                 # :237-272
        p = self.p
        V = len(unigram_scores)
        # sorted(range(V), reverse=True, key=score)[:k] (:240-241): a stable sort of the negated scores keeps the
        # lower token id first among ties, like reverse=True on a stable sort does
        us = np.asarray(unigram_scores, dtype=np.float64)
        order = np.argsort(-us, kind="stable")[:p["use_top_k_unigrams"]].tolist()
        kept = [t for t in order if t not in given]
        unigram_scores = us.tolist() if not isinstance(unigram_scores, list) else unigram_scores
        self.need_ranges([(t,) for t in kept])               # only the kept ones can score above zero
        table = [0.0] * V
        for t in kept:
            c = self.count((t,))
            if c == 0:
                continue
            if p["use_fm_index_frequency"]:
                sc = max(self.contrast(unigram_scores[t], c), 0.0)
            else:
                sc = max(unigram_scores[t] - cutoff, 0.0) ** p["alpha"]
            if sc != 0.0:
                table[t] = sc
        return table


def aggregate_evidence(ngrams_and_scores, unigram_scores: Optional[List[float]] = None, index=None,
                       max_occurrences_1: int = 1500, max_occurrences_2: int = 10_000_000,
                       n_docs_complete_score: int = 500, alpha: float = 2.0, beta: float = 0.8,
                       length_penalty: float = 0.0, use_fm_index_frequency: bool = True,
                       add_best_unigrams_to_ngrams: bool = False, use_top_k_unigrams=1000, sort_by_length=False,
                       sort_by_freq=False, smoothing=5.0, allow_overlaps=False, single_key=0.0,
                       single_key_add_unigrams=False, unigrams_ignore_free_places=False):
    """seal/keys.py:178-497.  Returns (results, all_ngrams): results = {doc: [score, [(key, score)...],
    None, doc_tokens, [best key, best score]]} sorted by descending score; all_ngrams = {key: score}."""
    ev = _Evidence(index, dict(alpha=alpha, beta=beta, length_penalty=length_penalty, smoothing=smoothing,
                               use_fm_index_frequency=use_fm_index_frequency, use_top_k_unigrams=use_top_k_unigrams))
    keys = [((k.tolist() if hasattr(k, "tolist") else list(k)), s) for k, s in ngrams_and_scores]
    cutoff = None
    if not use_fm_index_frequency:                                                          # :198-205
        cutoff = min(keys, key=lambda ks: ks[1])[1] - 0.1 if keys else [][0]
    ev.need_ranges([k for k, _ in keys])
    counts = {(): len(index)}                                                               # :196
    for k, _ in keys:
        counts[tuple(k)] = ev.count(k)
    given = {0, 1, 2} | {k[0] for k, _ in keys if len(k) == 1}                              # :207-211
    keys = [(k, ev.key_score(k, s, cutoff)) for k, s in keys]

    if unigram_scores is not None:
        unigram_scores = ev.unigram_table(unigram_scores, given, cutoff)
        if add_best_unigrams_to_ngrams:                                                     # :274-278
            extra = sorted(range(len(unigram_scores)), key=lambda t: -unigram_scores[t])[:len(keys)]
            ev.need_ranges([(t,) for t in extra])
            for t in extra:
                counts[(t,)] = ev.count((t,))
                keys.append(([t], unigram_scores[t]))

    # rare keys drive the first stage; frequent ones only take part in the full scoring  (:280-314)
    rare, freq = {}, {}
    for k, sc in keys:
        c = ev.count(k)
        if c > max_occurrences_2 or sc == 0.0:
            continue
        (freq if (c > max_occurrences_1 or sc < 0.0) else rare)[tuple(k)] = sc
    by_score = lambda kv: kv[1]
    rare = dict(sorted(rare.items(), key=by_score, reverse=True))
    freq = dict(sorted(freq.items(), key=by_score, reverse=True))
    all_ngrams = dict(sorted(list(rare.items()) + list(freq.items()), key=by_score, reverse=True))

    # -- batch 2: every SA row of every rare key, located and mapped to its document at once ------
    spans = []
    for k in rare:
        lo, hi = ev.range_of[k]
        spans.append((lo, max(lo, min(hi, lo + max_occurrences_1))))
    if any(b > a for a, b in spans):
        rows = np.concatenate([np.arange(a, b, dtype=np.uint64) for a, b in spans])
        pos, doc = index.locate_rows(rows)
    else:
        pos, doc = np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.int64)
    sort_mode = 1 if sort_by_length else (2 if sort_by_freq else 0)
    empty_count = int(counts[()])

    # first stage (:316-368) in native code: coverage of token positions, one credit per (key, document), damping
    # of repeated token types, shortlist of the n_docs_complete_score best documents
    rk = _FlatKeys(list(rare.items()), counts)
    span_off = np.zeros(len(spans) + 1, dtype=np.int64)
    np.cumsum([b - a for a, b in spans], out=span_off[1:])
    shortlist = np.zeros(max(len(pos), 1), dtype=np.int64); n_short = C.c_int64(0)
    _evcheck(lib.sealev_first_stage(len(rk), rk.tok.ctypes.data, rk.off.ctypes.data, rk.score.ctypes.data, rk.count.ctypes.data,
                                    empty_count, span_off.ctypes.data, np.ascontiguousarray(pos, dtype=np.uint64).ctypes.data,
                                    np.ascontiguousarray(doc, dtype=np.int64).ctypes.data, sort_mode, int(bool(allow_overlaps)),
                                    float(beta), float(single_key), int(n_docs_complete_score), shortlist.ctypes.data,
                                    C.byref(n_short)))
    shortlist = shortlist[:n_short.value].tolist()

    # -- batch 3: the shortlisted documents' tokens -----------------------------------------------
    fetch = getattr(index, "get_docs_arrays", None)
    texts = fetch(shortlist) if fetch else [np.asarray(t, dtype=np.int64) for t in index.get_docs(shortlist)]
    docs_arr = []
    for t in texts:                                                                         # :389: [2] + doc[:-1]
        a = np.empty(max(len(t), 1), dtype=np.int64)
        a[0] = 2; a[1:] = t[:-1]
        docs_arr.append(a)
    docs_tok = [a.tolist() for a in docs_arr]
    scored = [(k, v) for k, v in all_ngrams.items() if len(k) >= 1 and v > 0.0]             # trie contents, :378-385
    sk = _FlatKeys(scored, counts)
    doc_off = np.zeros(len(docs_tok) + 1, dtype=np.int64)
    np.cumsum([len(t) for t in docs_tok], out=doc_off[1:])
    flat = np.concatenate(docs_arr) if docs_arr else np.zeros(0, dtype=np.int64)
    uni = np.ascontiguousarray(unigram_scores, dtype=np.float64) if unigram_scores is not None else None
    n = len(docs_tok)
    out_score = np.zeros(max(n, 1)); out_best = np.zeros(max(n, 1), dtype=np.int64); out_best_score = np.zeros(max(n, 1))
    pick_off = np.zeros(n + 1, dtype=np.int64)
    # a document cannot pick more keys + unigram types than it has tokens -- when overlaps are forbidden.  With
    # allow_overlaps every nested / overlapping match is picked, so the buffer simply grows on SEALFM_ECAPACITY.
    cap = int(doc_off[-1]) * 2 + 16
    lib.sealev_set_sum_mode(1 if sys.version_info >= (3, 12) else 0)     # how this interpreter's sum() adds floats (:476)
    while True:
        pick_key = np.zeros(cap, dtype=np.int64); pick_score = np.zeros(cap)
        rc = lib.sealev_score_docs(len(sk), sk.tok.ctypes.data, sk.off.ctypes.data, sk.score.ctypes.data, sk.count.ctypes.data,
                                   empty_count, n, flat.ctypes.data, doc_off.ctypes.data,
                                   uni.ctypes.data if uni is not None else None, len(uni) if uni is not None else 0, sort_mode,
                                   int(bool(allow_overlaps)), int(bool(unigrams_ignore_free_places)),
                                   int(bool(single_key_add_unigrams)), float(beta), float(single_key), out_score.ctypes.data,
                                   out_best.ctypes.data, out_best_score.ctypes.data, pick_off.ctypes.data, pick_key.ctypes.data,
                                   pick_score.ctypes.data, cap)
        if rc == -6 and cap < (1 << 34):             # SEALFM_ECAPACITY
            cap *= 4
            continue
        _evcheck(rc)
        break
    results = {}
    pk, ps, po = pick_key.tolist(), pick_score.tolist(), pick_off.tolist()
    for i, d in enumerate(shortlist):
        picked = [((scored[k][0] if k >= 0 else (-1 - k,)), s) for k, s in zip(pk[po[i]:po[i + 1]], ps[po[i]:po[i + 1]])]
        b = int(out_best[i])
        best = [scored[b][0], float(out_best_score[i])] if b >= 0 else [[], 0.0]
        results[d] = [float(out_score[i]), picked, None, docs_tok[i], best]
    return dict(sorted(results.items(), key=lambda kv: -kv[1][0])), all_ngrams              # :496-497


class _FlatKeys:
    """(key tuple, score) pairs flattened for the native calls (include/sealev.h)."""

    def __init__(self, items, counts):
        self.off = np.zeros(len(items) + 1, dtype=np.int64)
        np.cumsum([len(k) for k, _ in items], out=self.off[1:])
        self.tok = np.fromiter((t for k, _ in items for t in k), dtype=np.int64, count=int(self.off[-1])) if items else np.zeros(0, dtype=np.int64)
        if len(self.tok) == 0:
            self.tok = np.zeros(1, dtype=np.int64)
        self.score = np.array([s for _, s in items], dtype=np.float64) if items else np.zeros(1)
        self.count = np.array([counts[tuple(k)] for k, _ in items], dtype=np.int64) if items else np.zeros(1, dtype=np.int64)
        self.n = len(items)

    def __len__(self):
        return self.n


def _evcheck(code):
    if code != 0:
        from ._lib import SealB200Error
        raise SealB200Error(code, lib.sealev_last_error().decode(errors="replace"))


_END = -1                                         # trie slot holding (key, score) of a complete key


def _build_trie(scored):
    root = {}
    for k, sc in scored.items():
        node = root
        for t in k:
            node = node.setdefault(t, {})
        node[_END] = (k, sc)
    return root


def _scan_keys(toks, root):
    """All occurrences of the scored keys in one document -> {key: [score, [(start, end)...]]}, keys in the
    order the reference's open-match list discovers them (it is popped from its END at every position,
    keys.py:400-409, so the visiting order of the live partial matches flips from one token to the next;
    only the order in which equal-scored keys are met depends on it).  Partial matches are (start, trie node)
    pairs, so a position costs one dict lookup per live match instead of a tuple slice and two set probes."""
    hits = {}
    live = []
    for i, t in enumerate(toks):
        keep = []
        node = root.get(t)                        # the match starting here is visited first
        if node is not None:
            keep.append((i, node))
            end = node.get(_END)
            if end is not None:
                hits.setdefault(end[0], [end[1], []])[1].append((i, i + 1))
        for a, node in reversed(live):
            node = node.get(t)
            if node is not None:
                keep.append((a, node))
                end = node.get(_END)
                if end is not None:
                    hits.setdefault(end[0], [end[1], []])[1].append((a, i + 1))
        live = keep
    return hits
