"""TEST INFRASTRUCTURE ONLY.  CPU/torch restatement of seal/keys.py:64-141 (rescore_keys) and
:145-176 (compute_unigram_scores) for the installed transformers (the reference file imports
`more_itertools`/`seal` and uses HF-4.13 private helpers, so it is restated, line references kept).
Pinned: tests/golden/make_keys_golden.py runs the unmodified reference functions (stub modules + a proxy
for the one private helper) beside these restatements -- |diff| = 0 -- and stores their outputs."""
import torch
from transformers.modeling_outputs import BaseModelOutput


def strip(seq, symbols_start, symbols_end):                     # keys.py:53-61
    i = 0
    while i < len(seq) and seq[i] in symbols_start:
        i += 1
    j = len(seq)
    while j > i and seq[j - 1] in symbols_end:
        j -= 1
    return seq[i:j]


@torch.inference_mode()
def rescore_keys_oracle(model, inputs, list_of_decoded, batch_size=100, length_penalty=0.0, prefix=[], strip_from_bos=[],
                        strip_from_eos=[]):
    cfg = model.config
    if inputs is None:
        batch_in = [[cfg.bos_token_id, cfg.eos_token_id]] * len(list_of_decoded)
    else:
        batch_in = list(inputs)
    list_of_decoded = [[x[1] if isinstance(x[0], float) else x for x in xx] for xx in list_of_decoded]
    maxlen = max(len(i) for i in batch_in)
    input_ids = torch.stack([torch.LongTensor(list(i) + [cfg.pad_token_id] * (maxlen - len(i))) for i in batch_in], 0)
    attention_mask = (input_ids != cfg.pad_token_id).long()
    enc = model.get_encoder()(input_ids=input_ids, attention_mask=attention_mask).last_hidden_state      # :84-85
    flat = [(idx, di) for idx, ddi in enumerate(list_of_decoded) for di in ddi]
    all_out = {i: [] for i in range(len(list_of_decoded))}
    for b0 in range(0, len(flat), batch_size):
        batch = flat[b0:b0 + batch_size]
        idxs, orig, dec = [], [], []
        for i, di in batch:
            stripped = [cfg.decoder_start_token_id] + list(prefix) + strip(list(di), strip_from_bos, strip_from_eos)
            idxs.append(i); orig.append(list(di)); dec.append(stripped)
        T = max(len(d) for d in dec)
        dec_ids = torch.stack([torch.LongTensor(d + [cfg.pad_token_id] * (T - len(d))) for d in dec], 0)
        logits = model(attention_mask=attention_mask[idxs], encoder_outputs=BaseModelOutput(last_hidden_state=enc[idxs]),
                       decoder_input_ids=dec_ids[:, :-1], use_cache=False).logits                           # :122-127
        logprobs = logits.float().log_softmax(-1)
        logprobs = torch.gather(logprobs, -1, dec_ids[:, 1:].unsqueeze(-1)).squeeze(-1)
        logprobs[dec_ids[:, 1:] < 2] = 0.0                                                                  # :132
        logprobs = logprobs[:, len(prefix):].sum(-1).tolist()
        for i, di, ll in zip(idxs, orig, logprobs):
            all_out[i].append((ll / (len(di) ** length_penalty), di))
    return [v for k, v in sorted(all_out.items())]


@torch.no_grad()
def compute_unigram_scores_oracle(model, inputs, temperature=1.0, prefix=[]):
    cfg = model.config
    batch_in = list(inputs)
    maxlen = max(len(i) for i in batch_in)
    input_ids = torch.stack([torch.LongTensor(list(i) + [cfg.pad_token_id] * (maxlen - len(i))) for i in batch_in], 0)
    attention_mask = (input_ids != cfg.pad_token_id).long()
    dec = torch.full((input_ids.shape[0], 1 + len(prefix)), cfg.decoder_start_token_id, dtype=torch.long)
    for i, t in enumerate(prefix, start=1):
        dec[:, i] = t
    logits = model(input_ids=input_ids, attention_mask=attention_mask, decoder_input_ids=dec, use_cache=False).logits[:, len(prefix)]
    if temperature != 1.0:
        logits = logits / temperature
    return logits.float().log_softmax(-1)


# ------------------------------------------------------------------------------------------------
# Evidence aggregation (seal/keys.py:178-497), restated sequentially on an OracleIndex.  Pinned
# against the reference function itself by tests/golden/make_keys_golden.py (which imports the
# unmodified seal/keys.py with stub `seal` / `more_itertools` modules) -> tests/golden/keys_golden.json.
# ------------------------------------------------------------------------------------------------
import math
from collections import Counter


def _damp(types, score, seen, beta):                            # keys.py:186-191 `repetition`
    if not seen:
        return score
    types = set(types)
    return (1.0 - beta + (beta * len(types.difference(seen)) / len(types))) * score


def _lm_vs_corpus(sr, count, ntokens, smoothing):               # keys.py:220-223 / :250-253
    snr = math.log((count + smoothing) / (ntokens + smoothing))
    return (sr + math.log(1 - math.exp(snr))) - (snr + math.log(1 - math.exp(sr)))


def aggregate_evidence_oracle(ngrams_and_scores, unigram_scores=None, index=None, max_occurrences_1=1500,
                              max_occurrences_2=10_000_000, n_docs_complete_score=500, alpha=2.0, beta=0.8,
                              length_penalty=0.0, use_fm_index_frequency=True, add_best_unigrams_to_ngrams=False,
                              use_top_k_unigrams=1000, sort_by_length=False, sort_by_freq=False, smoothing=5.0,
                              allow_overlaps=False, single_key=0.0, single_key_add_unigrams=False,
                              unigrams_ignore_free_places=False):
    ntokens = float(index.beginnings[-1])                                                  # :193
    keys = [(k.tolist() if isinstance(k, torch.Tensor) else k, s) for k, s in ngrams_and_scores]
    counts = {(): len(index)}                                                              # :196
    cutoff = None
    if not use_fm_index_frequency:                                                         # :198-205
        cutoff = sorted(keys, key=lambda x: x[1])[0][1] - 0.1
    given_unigrams = {0, 1, 2}                                                             # :207
    for pos, (k, sr) in enumerate(keys):                                                   # :208-235
        if len(k) == 1:
            given_unigrams.add(k[0])
        c = index.get_count(k)
        counts[tuple(k)] = c
        if c == 0:
            sc = 0.0
        elif use_fm_index_frequency:
            sr -= 1e-10
            sr *= (1.0 - length_penalty) ** (len(k) - 1.0)
            sc = max(_lm_vs_corpus(sr, c, ntokens, smoothing), 0.0)
            sc **= alpha
        else:
            sc = max(sr - cutoff, 0.0)
            sc *= (1.0 - length_penalty) ** (len(k) - 1.0)
            sc **= alpha
        keys[pos] = (k, sc)

    if unigram_scores is not None:                                                         # :237-281
        unigram_scores = unigram_scores[:]
        ranked = sorted(range(len(unigram_scores)), reverse=True, key=lambda i: unigram_scores[i])
        keep = set(ranked[:use_top_k_unigrams])
        unigram_scores = [s if i in keep else float("-inf") for i, s in enumerate(unigram_scores)]
        for t in range(len(unigram_scores)):
            if t in given_unigrams:
                unigram_scores[t] = 0.0
                continue
            sr = unigram_scores[t]
            c = index.get_count([t])
            if c == 0:
                sc = 0.0
            elif use_fm_index_frequency:
                sc = max(_lm_vs_corpus(sr, c, ntokens, smoothing), 0.0)
            else:
                sc = max(sr - cutoff, 0.0)
                sc **= alpha
            unigram_scores[t] = sc if sc != 0.0 else 0.0
        if add_best_unigrams_to_ngrams:
            for t in sorted(range(len(unigram_scores)), key=lambda x: -unigram_scores[x])[:len(keys)]:
                counts[(t,)] = index.get_count([t])
                keys.append(([t], unigram_scores[t]))

    rare, freq = {}, {}                                                                    # :283-303
    for k, sc in keys:
        c = index.get_count(k)
        if c > max_occurrences_2 or sc == 0.0:
            continue
        (freq if (c > max_occurrences_1 or sc < 0.0) else rare)[tuple(k)] = sc
    rare = dict(sorted(rare.items(), key=lambda kv: kv[1], reverse=True))                  # :306-314
    freq = dict(sorted(freq.items(), key=lambda kv: kv[1], reverse=True))
    all_ngrams = dict(sorted(list(rare.items()) + list(freq.items()), key=lambda kv: kv[1], reverse=True))

    covered = set()                                                                        # :316-351
    first = {}                                  # doc -> [score, [(key, score)...], [best key, best score]]

    def entry(d):
        if d not in first:
            first[d] = [0.0, [], [[], 0.0]]
        return first[d]

    for k, sc in rare.items():
        seen_docs = set()
        lo, hi = index.get_range(list(k))
        for row in list(range(lo, hi))[:max_occurrences_1]:
            end = index.locate(row)
            start = end - len(k)
            d = index.get_doc_index(end)
            fresh = all(p not in covered for p in range(start, end))
            e = entry(d)
            if sort_by_length:
                better = (len(k), sc) > (len(e[2][0]), e[2][1])
            elif sort_by_freq:
                better = (-counts[tuple(k)], sc) > (-counts[tuple(e[2][0])], e[2][1])
            else:
                better = sc > e[2][1]
            if better:
                e[2] = [k, sc]
            if fresh:
                covered.update(range(start, end))
            if (fresh or allow_overlaps) and d not in seen_docs:
                seen_docs.add(d)
                e[0] += sc
                e[1].append((k, sc))

    for d, e in first.items():                                                             # :353-365
        seen, total = set(), 0.0
        for j, (k, sc) in enumerate(e[1]):
            types = set(k)
            adj = _damp(types, sc, seen, beta)
            total += adj
            e[1][j] = [k, adj]
            seen |= types
        e[0] = total

    shortlist = sorted(first.items(), key=lambda kv: (1.0 - single_key) * (-kv[1][0]) + single_key * (-kv[1][2][1]))
    shortlist = shortlist[:n_docs_complete_score]                                          # :367-368

    scored = {k: v for k, v in all_ngrams.items() if len(k) >= 1 and v > 0.0}              # trie contents, :378-385
    prefixes = {k[:n] for k in scored for n in range(1, len(k) + 1)}
    results = {}
    for d, _ in shortlist:                                                                 # :387-491
        toks = [2] + index.get_doc(d)[:-1]
        best = [[], 0.0]
        type_scores = {t: (unigram_scores[t] if unigram_scores is not None else 0.0) for t in toks}
        found = {}                              # key -> [score, [(start, end)...]] in the reference's discovery order
        live = []                               # start offsets of partial matches; the reference pops its list from
        for i in range(len(toks)):              # the END each step (:400-409), so the order flips every position
            nxt = []
            for a in reversed(live + [i]):
                k = tuple(toks[a:i + 1])
                if k not in prefixes:
                    continue
                nxt.append(a)
                if k in scored:
                    found.setdefault(k, [scored[k], []])[1].append((a, i + 1))
            live = nxt
        queue = []
        for k, (s, places) in found.items():                                               # :413-432
            if sort_by_length:
                ahead = (-len(k), -s) < (-len(best[0]), -best[1])
            elif sort_by_freq:
                ahead = (counts[tuple(k)], -s) < (counts[tuple(best[0])], -best[1])
            else:
                ahead = -s < -best[1]
            queue.extend((-s, k, s, a, b) for a, b in places)
            if ahead:
                best = [k, s]
        queue.sort()                            # == heap pops: all pushes precede all pops, tuples are distinct
        seen, picked, prev = set(), [], None
        free = [True] * len(toks)
        for _, k, s, a, b in queue:                                                        # :434-470
            types = set(k)
            if prev == k:
                adj = picked[-1][1]
            elif not types:
                adj = 0.0
            else:
                adj = _damp(types, s, seen, beta)
            if adj <= 0.0:
                continue
            if not (allow_overlaps or all(free[a:b])):
                continue
            if prev == k:
                picked[-1] = (k, adj)
            else:
                prev = k
                seen |= types
                picked.append((k, adj))
            free[a:b] = [False] * (b - a)
        if unigrams_ignore_free_places:
            free = [True] * len(free)
        multi = sum(s for _, s in picked)
        uni = 0.0
        for t in Counter(t for t, f in zip(toks, free) if f):                              # :479-486
            s = type_scores[t]
            if s > 0.0:
                s = _damp((t,), s, seen, beta)
                if s != 0.0:
                    uni += s
                    picked.append(((t,), s))
        single = best[1] + (uni if single_key_add_unigrams else 0.0)
        multi += uni
        results[d] = [(1.0 - single_key) * multi + single_key * single, picked, None, toks, best]
    results = dict(sorted(results.items(), key=lambda kv: -kv[1][0]))                      # :496
    return results, all_ngrams
