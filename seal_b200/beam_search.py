"""Drop-in for ``seal.beam_search`` (/root/reference/seal/beam_search.py) on the B200 kernels.

* ``IndexBasedLogitsProcessor`` — same constructor, attributes and HF ``LogitsProcessor`` protocol
  (``__call__(input_ids, scores) -> scores + mask``, beam_search.py:33-140); the FM-index work and
  the mask run as CUDA kernels on the tensors' device, no ``.tolist()`` / H2D round trips.
* ``fm_index_generate`` — same signature and return value (beam_search.py:391-557), one beam group, no
  sampling: encoder, every decoder step, log-softmax, processors, FM-index constraint, top-k and
  BeamSearchScorerWithMemory all run inside libsealb200.so.  ``keep_history=True`` is the path SEALSearcher
  uses; ``keep_history=False`` (transformers' stock BeamSearchScorer, the signature's default) and
  ``transformers_output=True`` run the same kernels and replay the stock scorer over the per-step records.
* ``SealBartEngine`` — device copy of an HF ``BartForConditionalGeneration``'s weights.

No CPU path: CPU tensors are rejected.
"""
import ctypes as C
import math
import os
import weakref
from typing import List, Optional

import numpy as np

from ._lib import lib, check, vp, ProcessorCfg, BartConfig, DecParams
from .index import FMIndex, SHIFT

stopword_token_ids = [10, 41, 660, 5, 1941, 20, 7, 6]      # beam_search.py:22-31


def _torch():
    import torch
    return torch


def _occurring_mask(index, vocab, device=None):
    """uint32 bitmask of index.occurring_distinct (first-step rule, beam_search.py:73-77)."""
    cache = index.__dict__.setdefault("_occ_cache", {})
    key = (vocab, str(device))
    if key not in cache:
        words = np.zeros((vocab + 31) // 32, dtype=np.uint32)
        toks = np.asarray([t for t in index.occurring_distinct if 0 <= t < vocab], dtype=np.int64)
        np.bitwise_or.at(words, toks >> 5, (np.uint32(1) << (toks & 31).astype(np.uint32)))
        if device is None:
            cache[key] = words
        else:
            torch = _torch()
            cache[key] = torch.from_numpy(words.view(np.int32)).to(device)
    return cache[key]


class IndexBasedLogitsProcessor:
    """beam_search.py:33-140.  (Does not inherit transformers.LogitsProcessor so that importing the
    drop-in never depends on the installed transformers version; HF only duck-types `__call__`.)"""

    def __init__(self, index: FMIndex, num_beams: int, pad_token_id: int = 0, eos_token_id: int = 2,
                 force_decoding_from: Optional[List[int]] = None, stop_at_count: int = 0,
                 always_allow_eos: bool = False, forced_bos_token_id: Optional[int] = None):
        self.index = index
        self.pad_token_id = pad_token_id
        self.eos_token_id = eos_token_id
        self._num_beams = num_beams
        self.log_odds_weight = 0.0
        self.force_decoding_from = force_decoding_from
        self.force_decoding_second_token = None
        self.block_initial_stopwords = False
        self.stop_at_count = stop_at_count
        self.always_allow_eos = always_allow_eos
        self.forced_bos_token_id = forced_bos_token_id

    def __call__(self, input_ids, scores):
        torch = _torch()
        if not (scores.is_cuda and input_ids.is_cuda):
            raise RuntimeError("seal_b200.IndexBasedLogitsProcessor runs on CUDA tensors only (no CPU fallback)")
        if scores.dtype != torch.float32:
            raise TypeError("scores must be float32 (the reference decodes in fp32)")
        ids = input_ids.to(torch.int64).contiguous()
        sc = scores.contiguous()
        R, t = ids.shape
        V = sc.shape[-1]
        out = torch.empty_like(sc)
        force = self.force_decoding_from or []
        farr = (C.c_int64 * max(len(force), 1))(*force)
        cfg = ProcessorCfg(self._num_beams, self.pad_token_id, self.eos_token_id, int(self.stop_at_count),
                           int(bool(self.always_allow_eos)),
                           -1 if self.forced_bos_token_id is None else int(self.forced_bos_token_id),
                           len(force), farr, SHIFT)
        with torch.cuda.device(sc.device):
            occ = _occurring_mask(self.index, V, sc.device)
            check(lib.sealdec_apply_index_mask_d(self.index._dev(), torch.cuda.current_stream().cuda_stream,
                                                 C.byref(cfg), ids.data_ptr(), R, t, occ.data_ptr(),
                                                 sc.data_ptr(), out.data_ptr(), V, sc.stride(0)))
        return out


class SealBartEngine:
    """Device-resident BART weights + workspace (include/sealdec.h `sealbart_t`)."""

    def __init__(self, state_dict, config, device=0, gemm_mode=None):
        # gemm_mode: 5 = 3xFP16 on CTA pairs (cta_group::2; default; small problems use mode 3's split-K kernel), 3 = 3xFP16 with
        # one CTA per tile, 2 = 3xTF32 (fp32 range, the automatic fallback on fp16 overflow).  $SEALB200_GEMM overrides the default.
        if gemm_mode is None:
            gemm_mode = int(os.environ.get("SEALB200_GEMM", "5"))
        self.gemm_mode = int(gemm_mode)
        d = int(config.d_model)
        self.config = config
        self.device = int(device)
        cfg = BartConfig(int(config.vocab_size), d, int(config.encoder_layers), int(config.decoder_layers),
                         int(config.decoder_attention_heads), int(config.decoder_ffn_dim),
                         int(config.max_position_embeddings), int(bool(getattr(config, "scale_embedding", False))),
                         int(gemm_mode))
        if config.encoder_ffn_dim != config.decoder_ffn_dim or config.encoder_attention_heads != config.decoder_attention_heads:
            raise ValueError("encoder/decoder shapes must match (bart-large layout)")
        if getattr(config, "activation_function", "gelu") != "gelu":
            raise ValueError("only the exact-erf 'gelu' activation of bart-large is implemented")
        h = vp()
        check(lib.sealbart_create(C.byref(cfg), self.device, C.byref(h)))
        self._h = h.value
        for k, v in state_dict.items():
            if k.endswith("embed_tokens.weight") and "model.shared.weight" in state_dict:
                continue                                    # tied aliases of model.shared.weight
            if k == "lm_head.weight" and "model.shared.weight" in state_dict and v.data_ptr() == state_dict["model.shared.weight"].data_ptr():
                continue
            a = np.ascontiguousarray(v.detach().to("cpu").float().numpy())
            check(lib.sealbart_set_tensor(self._h, k.encode(), a.ctypes.data, a.size))
        check(lib.sealbart_finalize(self._h))

    @classmethod
    def from_hf(cls, model, device=None, gemm_mode=None):
        torch = _torch()
        if device is None:
            p = next(model.parameters())
            device = p.device.index if p.is_cuda else torch.cuda.current_device()
        return cls(model.state_dict(), model.config, device=device, gemm_mode=gemm_mode)

    def __del__(self):
        h = self.__dict__.get("_h")
        if h:
            lib.sealbart_free(h)
            self._h = None

    def device_bytes(self):
        return int(lib.sealbart_device_bytes(self._h))

    def debug_step_logits(self, input_ids, attention_mask, num_beams, decoder_input_ids):
        """Teacher-forced logits of the last decoder position for explicit decoder inputs [R,t]."""
        ids = np.ascontiguousarray(np.asarray(input_ids, dtype=np.int64))
        am = np.ascontiguousarray(np.asarray(attention_mask, dtype=np.int64))
        dec = np.ascontiguousarray(np.asarray(decoder_input_ids, dtype=np.int64))
        Q, S = ids.shape
        R, t = dec.shape
        assert R == Q * num_beams
        out = np.empty((R, int(self.config.vocab_size)), dtype=np.float32)
        check(lib.sealdec_debug_step_logits(self._h, ids.ctypes.data, am.ctypes.data, Q, S, num_beams,
                                            dec.ctypes.data, t, out.ctypes.data))
        return out

    def last_phase_us(self):
        a = (C.c_double * 5)()
        check(lib.sealdec_last_phase_us(self._h, a))
        return {"encoder": a[0], "decoder_layers": a[1], "lm_head": a[2], "select_expand": a[3], "total": a[4]}

    def profile_gemm(self, enable):
        """Enable/disable CUDA-event bracketing of every GEMM launch; returns the record so far."""
        us = C.c_double(0); n = C.c_int64(0); fl = C.c_double(0)
        check(lib.sealdec_profile_gemm(self._h, int(bool(enable)), C.byref(us), C.byref(n), C.byref(fl)))
        return {"total_us": us.value, "launches": n.value, "flops": fl.value}

    def last_launch_count(self):
        return int(lib.sealdec_last_launch_count(self._h))


_ENGINES = weakref.WeakKeyDictionary()


def _engine_for(model):
    if isinstance(model, SealBartEngine):
        return model
    eng = _ENGINES.get(model)
    if eng is None:
        eng = SealBartEngine.from_hf(model)
        _ENGINES[model] = eng
    return eng


def _make_params(cfg, num_beams, min_length, max_length, length_penalty, eos_token_id, force_decoding_from,
                 always_allow_eos, disable_fm_index, stop_at_count, forced_bos_token_id):
    force = list(force_decoding_from or [])
    farr = (C.c_int64 * max(len(force), 1))(*force)
    none = lambda x: -1 if x is None else int(x)
    p = DecParams(int(num_beams), int(min_length if min_length is not None else -1), int(max_length),
                  float(length_penalty), int(eos_token_id), int(cfg.pad_token_id), int(cfg.decoder_start_token_id),
                  none(cfg.eos_token_id), none(getattr(cfg, "forced_eos_token_id", None)), none(forced_bos_token_id),
                  int(stop_at_count), int(bool(always_allow_eos)), int(bool(disable_fm_index)), 1, len(force), farr, SHIFT)
    p._keepalive = farr
    return p


def generate_records(model, index, input_ids, attention_mask, min_length=3, max_length=25, length_penalty=1.0,
                     num_beams=3, eos_token_id=None, force_decoding_from=None, always_allow_eos=False,
                     disable_fm_index=False, stop_at_count=0, forced_bos_token_id="config", want_ranges=True):
    """The C-ABI call with HOST buffers (sealdec_generate): returns the packed hypothesis records
    (scores [Q,H] f32, lens [Q,H] i32, tokens [Q,H,T] i32, valid [Q,H] u8, lo/hi [Q,H] u64)."""
    eng = _engine_for(model)
    cfg = eng.config
    if forced_bos_token_id == "config":
        forced_bos_token_id = getattr(cfg, "forced_bos_token_id", None)
    if eos_token_id is None:
        eos_token_id = cfg.eos_token_id
    ids = np.ascontiguousarray(np.asarray(input_ids.cpu() if hasattr(input_ids, "cpu") else input_ids, dtype=np.int64))
    am = np.ascontiguousarray(np.asarray(attention_mask.cpu() if hasattr(attention_mask, "cpu") else attention_mask, dtype=np.int64))
    Q, S = ids.shape
    p = _make_params(cfg, num_beams, min_length, max_length, length_penalty, eos_token_id, force_decoding_from,
                     always_allow_eos, disable_fm_index, stop_at_count, forced_bos_token_id)
    H = int(lib.sealdec_hyps_per_query(C.byref(p)))
    T = int(max_length)
    scores = np.empty((Q, H), dtype=np.float32); lens = np.empty((Q, H), dtype=np.int32)
    toks = np.empty((Q, H, T), dtype=np.int32); valid = np.empty((Q, H), dtype=np.uint8)
    lo = np.zeros((Q, H), dtype=np.uint64) if want_ranges else None
    hi = np.zeros((Q, H), dtype=np.uint64) if want_ranges else None
    fm_h = None
    occ_ptr = None
    if not disable_fm_index:
        if index._device is None:
            index.to_device(eng.device)
        fm_h = index._dev()
        occ = _occurring_mask(index, int(cfg.vocab_size))
        occ_ptr = occ.ctypes.data
    check(lib.sealdec_generate(eng._h, fm_h, occ_ptr, C.byref(p), ids.ctypes.data, am.ctypes.data, Q, S,
                               scores.ctypes.data, lens.ctypes.data, toks.ctypes.data, valid.ctypes.data,
                               lo.ctypes.data if want_ranges else None, hi.ctypes.data if want_ranges else None))
    return {"scores": scores, "lens": lens, "tokens": toks, "valid": valid, "lo": lo, "hi": hi}


class DeviceRecords:
    """The hypothesis records of one generate call in ONE device buffer (sharding.RecordLayout): the decode
    kernels write into it, `sharding.gather_buffers` moves it (NCCL, device to device), `.host()` reads it."""

    def __init__(self, layout, device):
        torch = _torch()
        self.layout = layout
        self.buf = torch.zeros(layout.nbytes, dtype=torch.uint8, device=device)
        self._base = self.buf.data_ptr()

    def ptr(self, name):
        return self._base + self.layout.offsets[name][0]

    @property
    def err_ptr(self):
        return self._base + self.layout.err_offset

    def set_filled(self, n):
        torch = _torch()
        self.buf[:8] = torch.from_numpy(np.asarray([n], dtype=np.int64).view(np.uint8)).to(self.buf.device)

    def host(self):
        """Blocking read-back: dict of numpy arrays (first dim = the queries filled) + 'errors'."""
        from .sharding import merge_gathered
        return merge_gathered([self.buf], self.layout)


def _side_stream(device):
    """One non-default stream per device for the asynchronous entry points (CUDA graphs cannot be captured on the
    legacy default stream, which is what torch.cuda.current_stream() is unless the caller changed it)."""
    torch = _torch()
    cache = _side_stream.__dict__.setdefault("cache", {})
    key = torch.device(device).index if not isinstance(device, int) else device
    if key not in cache:
        cache[key] = torch.cuda.Stream(device=key)
    return cache[key]


def generate_records_device(model, index, input_ids_d, attention_mask_d, min_length=3, max_length=25, length_penalty=1.0,
                            num_beams=3, eos_token_id=None, force_decoding_from=None, always_allow_eos=False,
                            disable_fm_index=False, stop_at_count=0, forced_bos_token_id="config", out=None,
                            src_tokens=-1, stream=None):
    """sealdec_generate_dx on DEVICE tensors, asynchronous: input_ids / attention_mask are int64 CUDA tensors
    [Q, S]; the records land in `out` (a DeviceRecords, created if None) on `stream` (default: the current stream if
    it is not the legacy default stream, else a per-device side stream that first waits for the current one).
    `src_tokens`: number of non-zero mask entries if the caller knows it (right-padded masks) — then the call never
    touches the host; -1 = unknown.  Errors are flags inside the buffer (`out.host()["errors"]`, include/sealdec.h)."""
    torch = _torch()
    from .sharding import RecordLayout
    eng = _engine_for(model)
    cfg = eng.config
    if forced_bos_token_id == "config":
        forced_bos_token_id = getattr(cfg, "forced_bos_token_id", None)
    if eos_token_id is None:
        eos_token_id = cfg.eos_token_id
    if not (input_ids_d.is_cuda and attention_mask_d.is_cuda) or input_ids_d.dtype != torch.int64 or attention_mask_d.dtype != torch.int64:
        raise TypeError("generate_records_device takes int64 CUDA tensors (use generate_records for host arrays)")
    ids = input_ids_d.contiguous(); am = attention_mask_d.contiguous()
    Q, S = ids.shape
    p = _make_params(cfg, num_beams, min_length, max_length, length_penalty, eos_token_id, force_decoding_from,
                     always_allow_eos, disable_fm_index, stop_at_count, forced_bos_token_id)
    H = int(lib.sealdec_hyps_per_query(C.byref(p)))
    dev = ids.device
    if out is None:
        out = DeviceRecords(RecordLayout(Q, H, int(max_length)), dev)
    lay = out.layout
    if lay.H != H or lay.T != int(max_length) or lay.Q < Q:
        raise ValueError("record buffer layout does not fit this call")
    fm_h = None; occ_ptr = None
    if not disable_fm_index:
        if index._device is None:
            index.to_device(eng.device)
        fm_h = index._dev()
        occ = _occurring_mask(index, int(cfg.vocab_size), dev)
        occ_ptr = occ.data_ptr()
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream()
        if stream is None:
            stream = cur if cur.cuda_stream != 0 else _side_stream(dev)
        if stream.cuda_stream != cur.cuda_stream:
            stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            out.set_filled(Q)
            check(lib.sealdec_generate_dx(eng._h, fm_h, occ_ptr, C.byref(p), ids.data_ptr(), am.data_ptr(), Q, S,
                                          stream.cuda_stream, out.ptr("scores"), out.ptr("lens"), out.ptr("tokens"),
                                          out.ptr("valid"), out.ptr("lo"), out.ptr("hi"), out.err_ptr, int(src_tokens)))
        for t in (ids, am, out.buf):
            t.record_stream(stream)
        if stream.cuda_stream != cur.cuda_stream:
            cur.wait_stream(stream)
    return out


def sharded_generate_records(model, index, input_ids, attention_mask, group=None, dst=0, **kw):
    """N-GPU generate: this rank decodes its contiguous block of the batch (host arrays in, as SEALSearcher holds
    them), the records stay on the device and ONE NCCL gather brings every rank's buffer to `dst`
    (SURVEY.md section 8e).  Returns the full-batch record arrays on `dst`, None elsewhere."""
    torch = _torch()
    from .sharding import sharded_generate
    eng = _engine_for(model)
    dev = torch.device("cuda", eng.device)
    ids_np = np.ascontiguousarray(np.asarray(input_ids, dtype=np.int64)); am_np = np.ascontiguousarray(np.asarray(attention_mask, dtype=np.int64))
    cfg = eng.config
    p = _make_params(cfg, kw.get("num_beams", 3), kw.get("min_length", 3), kw.get("max_length", 25), kw.get("length_penalty", 1.0),
                     kw.get("eos_token_id") or cfg.eos_token_id, kw.get("force_decoding_from"), False, False, 0, None)
    H = int(lib.sealdec_hyps_per_query(C.byref(p)))

    def fill(ids_blk, am_blk, layout):
        # per-engine staging: the same device buffers (inputs and record buffer) serve every call of a shape, so that
        # small shards replay the captured CUDA graph of the call instead of launching ~1 900 kernels eagerly
        cache = eng.__dict__.setdefault("_io_cache", {})
        n = len(ids_blk)
        key = (layout.Q, layout.H, layout.T, n, ids_blk.shape[1] if n else 0)
        if key not in cache:
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[key] = (DeviceRecords(layout, dev),
                          torch.empty((max(n, 1), ids_np.shape[1]), dtype=torch.int64, device=dev),
                          torch.empty((max(n, 1), ids_np.shape[1]), dtype=torch.int64, device=dev))
        rec, ids_d, am_d = cache[key]
        if n:
            right_padded = bool((np.diff(am_blk != 0, axis=1) <= 0).all()) and bool((am_blk[:, 0] != 0).all())
            ids_d.copy_(torch.from_numpy(ids_blk)); am_d.copy_(torch.from_numpy(am_blk))
            generate_records_device(eng, index, ids_d, am_d, out=rec,
                                    src_tokens=int((am_blk != 0).sum()) if right_padded else -2, **kw)
        else:
            rec.set_filled(0)
        return rec.buf

    return sharded_generate(fill, ids_np, am_np, H, int(kw.get("max_length", 25)), group=group, dst=dst)


def records_to_output(rec, length_penalty):
    """beam_search.py:555: [(score * len**lp, tokens) for every recorded hyp with score > -inf].
    The arithmetic is vectorised (float64, like the reference's Python floats; len**lp comes from a table filled
    with Python's own pow so that every bit matches); only the final tuples are built in a Python loop."""
    scores, lens, toks = rec["scores"], rec["lens"], rec["tokens"]
    Q = scores.shape[0]
    if scores.size == 0:
        return [[] for _ in range(Q)]
    pen_table = np.array([float(n) ** length_penalty if n > 0 else 1.0 for n in range(int(lens.max()) + 1)], dtype=np.float64)
    pen = pen_table[lens]
    hyp = scores.astype(np.float64) / pen                             # BeamHypothesesWithMemory.add, :752-755
    final = hyp * pen
    valid = hyp > float("-inf")
    # one flat list of the kept tokens (prefixes of the kept hypotheses), sliced per hypothesis
    n_kept = np.where(valid, lens, 0)
    keep_tok = valid[:, :, None] & (np.arange(toks.shape[2])[None, None, :] < lens[:, :, None])
    flat = toks[keep_tok].tolist()
    ends = np.cumsum(n_kept.reshape(-1)[valid.reshape(-1)]).tolist()
    fs = final[valid].tolist()
    per_query = valid.sum(axis=1).tolist()
    out = []
    i = 0; start = 0
    for cnt in per_query:
        row = []
        for j in range(i, i + cnt):
            e = ends[j]
            row.append((fs[j], flat[start:e]))
            start = e
        out.append(row)
        i += cnt
    return out


def _replay_beam_search_scorer(rec, num_beams, length_penalty, eos_token_id, pad_token_id, max_length):
    """keep_history=False (seal/beam_search.py:505-515): the reference hands the loop transformers 4.13's stock
    `BeamSearchScorer` instead of BeamSearchScorerWithMemory.  Both choose the next beams the same way (the first
    num_beams non-EOS candidates of the top 2*num_beams), so the beams evolve identically until a query is `done`,
    after which the stock scorer freezes it (pads) and `finalize` skips it.  The device path therefore runs the very
    same kernels, and the stock scorer's bookkeeping -- BeamHypotheses.add / worst_score / is_done, process's
    "EOS only if ranked inside the top num_beams", finalize's best-num_beams selection with an appended EOS -- is
    replayed here, on the host, over the per-step candidate records (a few thousand scalar operations per query).
    Returns (beams per query [(score, tokens)] in BeamHypotheses order, sequences int64 [Q*num_beams, L],
    sequence_scores float32 [Q*num_beams]).  transformers 4.13 is not vendored: restated from its published algorithm."""
    scores, lens, toks = rec["scores"], rec["lens"], rec["tokens"]
    Q, H = scores.shape
    B, K = num_beams, 2 * num_beams
    n_steps = (H - B) // K
    all_beams, best, best_scores = [], [], []
    for q in range(Q):
        beams = []; worst = 1e9; done = False

        def add(tokens, sum_logprobs):
            nonlocal worst
            score = sum_logprobs / (len(tokens) ** length_penalty)
            if len(beams) < B or score > worst:
                beams.append((score, tokens))
                if len(beams) > B:
                    order = sorted((sc, i) for i, (sc, _) in enumerate(beams))
                    del beams[order[0][1]]
                    worst = order[1][0]
                else:
                    worst = min(score, worst)

        for st in range(n_steps):
            if done:
                break
            cur_len = st + 1
            base = st * K
            cand_s = scores[q, base:base + K].tolist()
            nb = 0
            for rank in range(K):
                h = base + rank
                tok = int(toks[q, h, cur_len])
                if tok == eos_token_id:
                    if rank < B:
                        add(toks[q, h, :cur_len].tolist(), cand_s[rank])
                else:
                    nb += 1
                if nb == B:
                    break
            if nb < B:
                raise ValueError(f"At most {B} tokens can be equal to `eos_token_id: {eos_token_id}`.")
            if len(beams) >= B:
                done = worst >= max(cand_s) / cur_len ** length_penalty
        if not done:
            fb = n_steps * K
            for j in range(B):
                add(toks[q, fb + j, :lens[q, fb + j]].tolist(), float(scores[q, fb + j]))
        all_beams.append(list(beams))
        order = sorted(beams, key=lambda x: x[0])
        for _ in range(B):
            sc, t = order.pop()
            best.append(t); best_scores.append(sc)
    L = min(max(len(t) for t in best) + 1, max_length) if best else 0
    seq = np.full((len(best), L), pad_token_id, dtype=np.int64)
    for i, t in enumerate(best):
        seq[i, :len(t)] = t
        if len(t) < max_length:
            seq[i, len(t)] = eos_token_id
    return all_beams, seq, np.asarray(best_scores, dtype=np.float32)


def fm_index_generate(model, index: FMIndex, input_ids, attention_mask, min_length: int = 3, max_length: int = 25,
                      length_penalty: float = 1.0, num_beams: int = 3, diverse_bs_groups: int = 1,
                      diverse_bs_penalty: float = 0.0, eos_token_id: Optional[int] = None,
                      force_decoding_from: Optional[List[int]] = None, always_allow_eos: bool = False,
                      keep_history: bool = False, disable_fm_index: bool = False, sample: bool = False,
                      stop_at_count: int = 0, topk: int = 0, transformers_output: bool = False, **kwargs):
    """beam_search.py:391-557.  `model` is an HF BartForConditionalGeneration (its weights are
    mirrored on the GPU once and cached) or a SealBartEngine."""
    if diverse_bs_groups != 1 or sample or topk:
        raise NotImplementedError("diverse beam groups / sampling / top-k warping are outside the constrained-decoding hot path")
    forced_bos = kwargs.pop("forced_bos_token_id", "config")                     # :415-418
    torch = _torch()
    if keep_history:
        rec = generate_records(model, index, input_ids, attention_mask, min_length, max_length, length_penalty,
                               num_beams, eos_token_id, force_decoding_from, always_allow_eos, disable_fm_index,
                               stop_at_count, forced_bos, want_ranges=False)
        if transformers_output:
            # BeamSearchScorerWithMemory.finalize returns an UNINITIALISED [Q*num_beams, 3] tensor as `sequences`
            # (:727); zeros of that shape here
            return torch.zeros((rec["scores"].shape[0] * num_beams, 3), dtype=torch.long,
                               device=input_ids.device if hasattr(input_ids, "device") else "cpu")
        return records_to_output(rec, length_penalty)
    # ---- keep_history=False: transformers' stock BeamSearchScorer (:505-515), the reference's default --------------
    eng = _engine_for(model)
    cfg = eng.config
    eos = cfg.eos_token_id if eos_token_id is None else eos_token_id
    dev = torch.device("cuda", eng.device)
    ids_np = np.ascontiguousarray(np.asarray(input_ids.cpu() if hasattr(input_ids, "cpu") else input_ids, dtype=np.int64))
    am_np = np.ascontiguousarray(np.asarray(attention_mask.cpu() if hasattr(attention_mask, "cpu") else attention_mask, dtype=np.int64))
    out = generate_records_device(eng, index, torch.from_numpy(ids_np).to(dev), torch.from_numpy(am_np).to(dev), min_length,
                                  max_length, length_penalty, num_beams, eos_token_id, force_decoding_from, always_allow_eos,
                                  disable_fm_index, stop_at_count, forced_bos, src_tokens=-2)
    rec = out.host()
    if rec["errors"][1] and eng.gemm_mode >= 3:        # fp16 range exceeded: redo with the 3xTF32 kernels (sealdec.h)
        check(lib.sealbart_set_option(eng._h, b"gemm_mode", 2))
        try:
            rec = generate_records_device(eng, index, torch.from_numpy(ids_np).to(dev), torch.from_numpy(am_np).to(dev), min_length,
                                          max_length, length_penalty, num_beams, eos_token_id, force_decoding_from, always_allow_eos,
                                          disable_fm_index, stop_at_count, forced_bos, src_tokens=-2).host()
        finally:
            check(lib.sealbart_set_option(eng._h, b"gemm_mode", eng.gemm_mode))
    # the device's "fewer than num_beams non-EOS candidates" flag also fires for queries the stock scorer had already
    # frozen; the replay re-derives the condition per query and step and raises exactly where the reference does
    beams, seq, _ = _replay_beam_search_scorer(rec, num_beams, length_penalty, eos, cfg.pad_token_id, int(max_length))
    if transformers_output:
        return torch.from_numpy(seq).to(input_ids.device if hasattr(input_ids, "device") else "cpu")     # :388
    return [[(sc * (len(t) ** length_penalty), t) for sc, t in b if sc > float("-inf")] for b in beams]    # :555
