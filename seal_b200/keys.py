"""Drop-in for the decoder-side helpers of ``seal.keys`` (/root/reference/seal/keys.py) that run
right after every generate pass: ``rescore_keys`` (:64-141) and
``compute_unigram_scores`` (:145-176).  Same signatures and return values; the teacher-forced BART
pass runs on the kernels behind ``sealdec_teacher_forced`` (include/sealdec.h).  No CPU path."""
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
