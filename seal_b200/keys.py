"""Drop-in for ``seal.keys`` (/root/reference/seal/keys.py): the decoder-side helpers that run right
after every generate pass -- ``rescore_keys`` (:64-141) and
``compute_unigram_scores`` (:145-176), on the kernels behind ``sealdec_teacher_forced``
(include/sealdec.h) -- and the evidence aggregation (§8f rank 2) ``aggregate_evidence`` (:178-497),
whose FM-index accesses are three batched GPU launches.  Same signatures and return values.
No CPU path."""
import ctypes as C
from typing import List, Optional

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
    lp = lp.astype(np.float64)
    lp[dec[:, 1:] < 2] = 0.0                                                     # :132
    lp = lp[:, len(prefix):].sum(-1)                                             # :133-134
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
        order = sorted(range(V), reverse=True, key=lambda t: unigram_scores[t])
        kept = [t for t in order[:p["use_top_k_unigrams"]] if t not in given]
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
        spans.append((lo, min(hi, lo + max_occurrences_1)))
    if spans:
        rows = np.concatenate([np.arange(a, b, dtype=np.uint64) for a, b in spans]) if any(b > a for a, b in spans) \
            else np.zeros(0, dtype=np.uint64)
        pos, doc = index.locate_rows(rows) if len(rows) else (rows, rows.astype(np.int64))
        pos, doc = pos.tolist(), doc.tolist()
    else:
        pos, doc = [], []

    covered = set()                                                                         # :316-351
    stage1 = {}                                   # doc -> [sum, [(key, score)...], [best key, best score]]
    at = 0
    for (k, sc), (a, b) in zip(rare.items(), spans):
        n = len(k)
        rank_new = (n, sc) if sort_by_length else ((-counts[k], sc) if sort_by_freq else sc)
        credited = set()
        for j in range(at, at + max(b - a, 0)):
            end, d = pos[j], doc[j]
            e = stage1.get(d)
            if e is None:
                e = stage1[d] = [0.0, [], [[], 0.0]]
            bk, bs = e[2]
            rank_old = (len(bk), bs) if sort_by_length else ((-counts[tuple(bk)], bs) if sort_by_freq else bs)
            if rank_new > rank_old:
                e[2] = [k, sc]
            fresh = covered.isdisjoint(range(end - n, end))
            if fresh:
                covered.update(range(end - n, end))
            if (fresh or allow_overlaps) and d not in credited:
                credited.add(d)
                e[0] += sc
                e[1].append((k, sc))
        at += max(b - a, 0)

    for e in stage1.values():                                                               # :353-365
        seen, total = set(), 0.0
        for j, (k, sc) in enumerate(e[1]):
            types = set(k)
            adj = ev.damp(types, sc, seen)
            total += adj
            e[1][j] = [k, adj]
            seen |= types
        e[0] = total

    shortlist = sorted(stage1.items(), key=lambda kv: (1.0 - single_key) * (-kv[1][0]) + single_key * (-kv[1][2][1]))
    shortlist = [d for d, _ in shortlist[:n_docs_complete_score]]                           # :367-368

    # -- batch 3: the shortlisted documents' tokens -----------------------------------------------
    texts = index.get_docs(shortlist)
    scored = {k: v for k, v in all_ngrams.items() if len(k) >= 1 and v > 0.0}               # :378-385
    stems = {k[:n] for k in scored for n in range(1, len(k) + 1)}

    results = {}
    for d, text in zip(shortlist, texts):                                                   # :387-491
        toks = [2] + text[:-1]
        hits = _scan_keys(toks, scored, stems)
        best = [[], 0.0]
        queue = []
        for k, (s, places) in hits.items():                                                 # :413-432
            if sort_by_length:
                ahead = (-len(k), -s) < (-len(best[0]), -best[1])
            elif sort_by_freq:
                ahead = (counts[k], -s) < (counts[tuple(best[0])], -best[1])
            else:
                ahead = -s < -best[1]
            if ahead:
                best = [k, s]
            queue += [(-s, k, s, a, b) for a, b in places]
        queue.sort()                              # the reference's heap is filled completely before it is drained
        seen, picked, prev = set(), [], None
        free = [True] * len(toks)
        for _, k, s, a, b in queue:                                                         # :434-470
            if prev == k:
                adj = picked[-1][1]
            else:
                adj = ev.damp(k, s, seen)
            if adj <= 0.0 or not (allow_overlaps or all(free[a:b])):
                continue
            if prev != k:
                prev = k
                seen.update(k)
                picked.append((k, adj))
            free[a:b] = [False] * (b - a)
        if unigrams_ignore_free_places:
            free = [True] * len(toks)
        total = sum(s for _, s in picked)
        uni = 0.0
        if unigram_scores is not None:                                                      # :479-486 (all-zero otherwise)
            for t in _Counter(t for t, f in zip(toks, free) if f):
                s = unigram_scores[t]
                if s > 0.0:
                    s = ev.damp((t,), s, seen)
                    if s != 0.0:
                        uni += s
                        picked.append(((t,), s))
        lone = best[1] + (uni if single_key_add_unigrams else 0.0)
        total += uni
        results[d] = [(1.0 - single_key) * total + single_key * lone, picked, None, toks, best]
    return dict(sorted(results.items(), key=lambda kv: -kv[1][0])), all_ngrams              # :496-497


def _scan_keys(toks, scored, stems):
    """All occurrences of the scored keys in one document -> {key: [score, [(start, end)...]]}, keys in the
    order the reference's open-match list discovers them (it is popped from its END at every position,
    keys.py:400-409, so the visiting order of the live partial matches flips from one token to the next;
    only the order in which equal-scored keys are met depends on it)."""
    hits = {}
    live = []
    for i in range(len(toks)):
        keep = []
        for a in reversed(live + [i]):
            k = tuple(toks[a:i + 1])
            if k in stems:
                keep.append(a)
                if k in scored:
                    hits.setdefault(k, [scored[k], []])[1].append((a, i + 1))
        live = keep
    return hits
