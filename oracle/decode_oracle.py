"""TEST INFRASTRUCTURE ONLY — never imported by the product package (seal_b200/).

CPU/torch restatement of the reference's constrained beam search, /root/reference/seal/
beam_search.py, for the transformers version installed here (5.5; the reference pins 4.13 whose
private generation helpers — `_get_logits_processor`, `BeamScorer`, `_reorder_cache`, ... — no
longer exist, so seal/beam_search.py cannot be imported).  Each block cites the
reference lines it follows.  The model is any callable

    step_logits(decoder_input_ids[R,t]) -> logits[R,V] (fp32, last position)

so the same loop runs on HF BART (``HFBartStepper`` below: eager fp32 PyTorch, the reference's
own arithmetic for the decoder, transformers being the un-vendored dependency that owns it) and
on synthetic logit tables in tests.

Parity status: PINNED against the reference's own decode code.  tests/golden/make_decode_golden.py
imports /root/reference/seal/beam_search.py unmodified (stub modules for the transformers-4.13 names
that no longer exist, a `Bart413Adapter` for the private GenerationMixin helpers) and runs its
fm_index_generate / constrained_beam_search / IndexBasedLogitsProcessor / BeamSearchScorerWithMemory
beside this restatement: identical hypothesis lists in identical order, |dscore| = 0, for nine parameter
sets; the reference's outputs are the fixture tests/golden/decode_golden.json.  What remains an
assumption is transformers 4.13 itself (not installed): the processor list of SURVEY.md §H3 (kept
switchable) and InfNanRemove's 4.13 form (NaN/+inf only).
"""
import math
from typing import Callable, List, Optional, Sequence

import torch

NEG_INF = float("-inf")


# ---- transformers-4.13 logits processors the reference obtains from model._get_logits_processor
# (seal/beam_search.py:430-445), SURVEY.md §H3.  In-place on `scores`, like 4.13. -------------------
def proc_min_length(input_ids, scores, min_length, eos_token_id):
    if input_ids.shape[-1] < min_length:
        scores[:, eos_token_id] = NEG_INF
    return scores


def proc_forced_bos(input_ids, scores, bos_token_id):
    if input_ids.shape[-1] == 1:
        keep = scores.new_full(scores.shape, NEG_INF)
        keep[:, bos_token_id] = 0
        return keep
    return scores


def proc_forced_eos(input_ids, scores, max_length, eos_token_id):
    if input_ids.shape[-1] == max_length - 1:
        keep = scores.new_full(scores.shape, NEG_INF)
        keep[:, eos_token_id] = 0
        return keep
    return scores


def proc_inf_nan(input_ids, scores):
    scores[scores != scores] = 0.0
    scores[scores == float("inf")] = torch.finfo(scores.dtype).max
    return scores


class IndexBasedLogitsProcessorOracle:
    """seal/beam_search.py:33-140, line by line, on an oracle index (oracle.fm_oracle.OracleIndex)."""

    def __init__(self, index, num_beams, pad_token_id=0, eos_token_id=2, force_decoding_from=None,
                 stop_at_count=0, always_allow_eos=False, forced_bos_token_id=None):
        self.index = index
        self.pad_token_id = pad_token_id
        self.eos_token_id = eos_token_id
        self._num_beams = num_beams
        self.force_decoding_from = force_decoding_from
        self.stop_at_count = stop_at_count
        self.always_allow_eos = always_allow_eos
        self.forced_bos_token_id = forced_bos_token_id

    def __call__(self, input_ids, scores):
        mask = torch.full_like(scores, NEG_INF)                                   # :64
        if self.forced_bos_token_id is not None:                                  # :66-71
            if input_ids.size(1) == 1:
                mask[:, self.forced_bos_token_id] = 0.0
                return scores + mask
            input_ids = input_ids[:, 1:]
        if input_ids.size(1) == 1:                                                # :73-77
            distinct = torch.LongTensor(self.index.occurring_distinct)
            mask[:, distinct] = 0.0
        else:
            ids = input_ids.view(-1, self._num_beams, input_ids.shape[-1]).tolist()  # :81
            lows, highs, counts = [], [], []
            for beam_sent in ids:                                                 # :87-105
                for sent in beam_sent:
                    if sent[-1] in (self.eos_token_id, self.pad_token_id):
                        low = high = count = 0
                    elif self.force_decoding_from is not None:
                        low, high = self.index.get_range(self.force_decoding_from + sent[1:])
                        count = self.index.get_count(self.force_decoding_from + sent[1:-1])
                    else:
                        low, high = self.index.get_range(sent[1:])
                        count = self.index.get_count(sent[1:-1])
                    lows.append(low); highs.append(high); counts.append(count)
            results = self.index.get_distinct_count_multi(lows, highs)            # :107
            k = 0
            for batch_id, beam_sent in enumerate(ids):                            # :111-135
                for beam_id, sent in enumerate(beam_sent):
                    if self.stop_at_count > 0 and counts[k] <= self.stop_at_count:
                        distinct = [self.eos_token_id]
                    elif sent[-1] == self.eos_token_id:
                        distinct = [self.pad_token_id]
                    elif sent[-1] == self.pad_token_id:
                        distinct = [self.pad_token_id]
                    else:
                        distinct = results[k][0]
                    k += 1
                    mask[batch_id * self._num_beams + beam_id, torch.LongTensor(distinct)] = 0
        if self.always_allow_eos:                                                 # :137-138
            mask[:, self.eos_token_id] = 0.0
        return scores + mask                                                      # :140


class HypsWithMemory:
    """BeamHypothesesWithMemory (:737-758) + the part of BeamSearchScorerWithMemory.process that
    touches it (:642-695)."""

    def __init__(self, length_penalty, max_length):
        self.length_penalty = length_penalty
        self.max_length = max_length
        self.beams = []         # (score, tokens list, constrained_score)

    def add(self, tokens, sum_logprobs, constrained):
        score = sum_logprobs / (len(tokens) ** self.length_penalty)               # :752-755
        self.beams.append((score, list(tokens), constrained))


class HFBeamHypotheses413:
    """transformers 4.13 `BeamHypotheses` (generation_beam_search.py), the container behind `BeamSearchScorer` that
    fm_index_generate builds when keep_history=False (seal/beam_search.py:505-515).  transformers 4.13 is an
    un-vendored dependency (requirements.txt:6) and the class no longer exists in the installed 5.5, so this is a
    restatement of its PUBLISHED algorithm -- PARITY UNPINNED for this class and `HFBeamSearchScorer413` (no
    reference-run fixture can be produced here); everything else on the path stays pinned."""

    def __init__(self, num_beams, length_penalty, early_stopping):
        self.length_penalty = length_penalty
        self.early_stopping = early_stopping
        self.num_beams = num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / (len(hyp) ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, list(hyp)))
            if len(self) > self.num_beams:
                sorted_next_scores = sorted([(s, idx) for idx, (s, _) in enumerate(self.beams)])
                del self.beams[sorted_next_scores[0][1]]
                self.worst_score = sorted_next_scores[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


class HFBeamSearchScorer413:
    """transformers 4.13 `BeamSearchScorer.process / finalize` for one beam group (see HFBeamHypotheses413 for the
    parity status).  `process` returns (scores, tokens, indices) of the next beams as flat lists."""

    def __init__(self, batch_size, num_beams, length_penalty, do_early_stopping=False, num_beam_hyps_to_keep=1):
        self.num_beams = num_beams
        self.num_beam_hyps_to_keep = num_beam_hyps_to_keep
        self._beam_hyps = [HFBeamHypotheses413(num_beams, length_penalty, do_early_stopping) for _ in range(batch_size)]
        self._done = [False] * batch_size

    @property
    def is_done(self):
        return all(self._done)

    def process(self, input_ids, next_scores, next_tokens, next_indices, pad_token_id, eos_token_id):
        cur_len = input_ids.shape[-1]
        B = self.num_beams
        nb_scores, nb_tokens, nb_idx = [], [], []
        for b, hyp in enumerate(self._beam_hyps):
            if self._done[b]:                                   # pad the batch
                nb_scores += [0.0] * B; nb_tokens += [pad_token_id] * B; nb_idx += [0] * B
                continue
            beam_idx = 0
            sc, tk, ix = [], [], []
            for rank, (tok, s, bi) in enumerate(zip(next_tokens[b].tolist(), next_scores[b].tolist(), next_indices[b].tolist())):
                bbi = b * B + bi
                if eos_token_id is not None and tok == eos_token_id:
                    if rank >= B:                               # not among the top num_beams: never added
                        continue
                    hyp.add(input_ids[bbi].tolist(), s)
                else:
                    sc.append(s); tk.append(tok); ix.append(bbi)
                    beam_idx += 1
                if beam_idx == B:
                    break
            if beam_idx < B:
                raise ValueError(f"At most {B} tokens in {next_tokens[b]} can be equal to `eos_token_id: {eos_token_id}`. "
                                 "Make sure {next_tokens[batch_idx]} are corrected.")
            nb_scores += sc; nb_tokens += tk; nb_idx += ix
            self._done[b] = self._done[b] or hyp.is_done(max(next_scores[b].tolist()), cur_len)
        return nb_scores, nb_tokens, nb_idx

    def finalize(self, input_ids, final_beam_scores, max_length, pad_token_id, eos_token_id):
        B = self.num_beams
        for b, hyp in enumerate(self._beam_hyps):
            if self._done[b]:
                continue
            for beam_id in range(B):
                hyp.add(input_ids[b * B + beam_id].tolist(), float(final_beam_scores[b * B + beam_id]))
        best, best_scores = [], []
        for hyp in self._beam_hyps:
            sorted_hyps = sorted(hyp.beams, key=lambda x: x[0])
            for _ in range(self.num_beam_hyps_to_keep):
                s, t = sorted_hyps.pop()
                best.append(t); best_scores.append(s)
        lens = [len(t) for t in best]
        sent_max_len = min(max(lens) + 1, max_length)
        decoded = torch.full((len(best), sent_max_len), pad_token_id, dtype=torch.long)
        for i, t in enumerate(best):
            decoded[i, :lens[i]] = torch.tensor(t, dtype=torch.long)
            if lens[i] < max_length:
                decoded[i, lens[i]] = eos_token_id
        return decoded, torch.tensor(best_scores, dtype=torch.float32)


def constrained_beam_search_oracle(
        step_logits: Callable[[torch.Tensor], torch.Tensor],
        batch_size: int,
        index,
        num_beams: int,
        min_length: int,
        max_length: int,
        length_penalty: float = 1.0,
        eos_token_id: int = 2,
        pad_token_id: int = 1,
        decoder_start_token_id: int = 2,
        model_eos_token_id: int = 2,
        forced_eos_token_id: Optional[int] = 2,
        forced_bos_token_id: Optional[int] = None,
        force_decoding_from: Optional[List[int]] = None,
        stop_at_count: int = 0,
        always_allow_eos: bool = False,
        disable_fm_index: bool = False,
        processors: Sequence[str] = ("min_length", "forced_bos", "forced_eos", "inf_nan"),
        reorder: Optional[Callable[[torch.Tensor], None]] = None,
        trace: Optional[list] = None,
        keep_history: bool = True,
        transformers_output: bool = False):
    """fm_index_generate (:391-557, keep_history=True, diverse_bs_groups=1, sample=False, topk=0)
    + constrained_beam_search (:143-389) + BeamSearchScorerWithMemory (:559-735).

    Returns per query a list of (score * len**lp, tokens, constrained_score) for every recorded
    hypothesis with score > -inf (:555); `constrained_score` is extra bookkeeping (the top-k value
    on the constrained tensor, -inf for the tie-filled picks of SURVEY.md §H4).
    """
    cdp = None
    if not disable_fm_index:                                                      # :457-467
        cdp = IndexBasedLogitsProcessorOracle(index, num_beams, pad_token_id=pad_token_id,
                                              eos_token_id=eos_token_id or model_eos_token_id,
                                              force_decoding_from=force_decoding_from, stop_at_count=stop_at_count,
                                              always_allow_eos=always_allow_eos,
                                              forced_bos_token_id=forced_bos_token_id)
    hyps = [HypsWithMemory(length_penalty, max_length) for _ in range(batch_size)]   # :493-503
    hf_scorer = None
    if not keep_history:                                                          # :505-515
        hf_scorer = HFBeamSearchScorer413(batch_size, num_beams, length_penalty, do_early_stopping=False,
                                          num_beam_hyps_to_keep=num_beams)
    R = batch_size * num_beams
    input_ids = torch.full((R, 1), decoder_start_token_id, dtype=torch.long)      # :485-489, :517-521
    beam_scores = torch.zeros((batch_size, num_beams), dtype=torch.float)         # :214-216
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    while True:                                                                   # :219
        logits = step_logits(input_ids).float()                                   # :231-244
        scores = torch.log_softmax(logits, dim=-1)                                # :251-253
        cur = input_ids
        for p in processors:                                                      # :255, §H3 order
            if p == "min_length" and min_length is not None and min_length > -1:
                scores = proc_min_length(cur, scores, min_length, model_eos_token_id)
            elif p == "forced_bos" and forced_bos_token_id is not None:
                scores = proc_forced_bos(cur, scores, forced_bos_token_id)
            elif p == "forced_eos" and forced_eos_token_id is not None:
                scores = proc_forced_eos(cur, scores, max_length, forced_eos_token_id)
            elif p == "inf_nan":
                scores = proc_inf_nan(cur, scores)
        processed = scores
        next_scores = processed + beam_scores[:, None]                            # :258
        if cdp is not None:                                                       # :260-262
            constrained = cdp(input_ids, processed) + beam_scores[:, None]
        else:
            constrained = next_scores
        V = next_scores.shape[-1]
        next_scores = next_scores.view(batch_size, num_beams * V)                 # :302-307
        constrained = constrained.view(batch_size, num_beams * V)
        top_c, top_i = torch.topk(constrained, 2 * num_beams, dim=1, largest=True, sorted=True)
        top_s = next_scores.gather(-1, top_i)
        next_indices = torch.div(top_i, V, rounding_mode="floor")                 # :309-310 (§H6)
        next_tokens = top_i % V
        if trace is not None:
            trace.append({"input_ids": input_ids.clone(), "beam_scores": beam_scores.clone(),
                          "top_scores": top_s.clone(), "top_constrained": top_c.clone(),
                          "top_tokens": next_tokens.clone(), "top_beams": next_indices.clone()})
        cur_len = input_ids.shape[-1]
        if hf_scorer is not None:                                                 # BeamSearchScorer.process (transformers 4.13)
            sc_l, tk_l, ix_l = hf_scorer.process(input_ids, top_s, next_tokens, next_indices, pad_token_id, eos_token_id)
            beam_scores = torch.tensor(sc_l, dtype=torch.float)
            beam_idx_flat = torch.tensor(ix_l, dtype=torch.long)
            input_ids = torch.cat([input_ids[beam_idx_flat, :], torch.tensor(tk_l, dtype=torch.long).view(-1, 1)], dim=-1)
            if reorder is not None:
                reorder(beam_idx_flat)
            if hf_scorer.is_done or input_ids.shape[-1] >= max_length:            # :340
                break
            continue
        # BeamSearchScorerWithMemory.process (:614-703)
        nb_scores = torch.zeros((batch_size, num_beams)); nb_tokens = torch.zeros((batch_size, num_beams), dtype=torch.long)
        nb_idx = torch.zeros((batch_size, num_beams), dtype=torch.long)
        for b in range(batch_size):
            beam_idx = 0
            broken = False
            for tok, sc, bi, cs in zip(next_tokens[b].tolist(), top_s[b].tolist(), next_indices[b].tolist(), top_c[b].tolist()):
                bbi = b * num_beams + bi
                hyps[b].add(input_ids[bbi].tolist() + [tok], sc, cs)              # :662-668
                if broken:
                    pass
                elif eos_token_id is not None and tok == eos_token_id:            # :673-674
                    pass
                else:
                    nb_scores[b, beam_idx] = sc; nb_tokens[b, beam_idx] = tok; nb_idx[b, beam_idx] = bbi
                    beam_idx += 1
                if beam_idx == num_beams:
                    broken = True
            if beam_idx < num_beams:                                              # :687-690
                raise ValueError(f"At most {num_beams} tokens can be equal to `eos_token_id: {eos_token_id}`.")
        beam_scores = nb_scores.view(-1)
        beam_idx_flat = nb_idx.view(-1)
        input_ids = torch.cat([input_ids[beam_idx_flat, :], nb_tokens.view(-1, 1)], dim=-1)   # :326
        if reorder is not None:
            reorder(beam_idx_flat)                                                # :331-332
        if cur_len + 0 >= max_length or input_ids.shape[-1] >= max_length:        # :340, :757-758, MaxLengthCriteria
            break
    if hf_scorer is not None:
        sequences, seq_scores = hf_scorer.finalize(input_ids, beam_scores, max_length, pad_token_id, eos_token_id)   # :342-350
        if transformers_output:
            return sequences                                                      # :388 (return_dict_in_generate is False)
        return [[(s * (len(t) ** length_penalty), t, float("nan")) for (s, t) in h.beams if s > NEG_INF]
                for h in hf_scorer._beam_hyps]                                    # :555
    if trace is not None:
        trace.append({"final_input_ids": input_ids.clone(), "final_beam_scores": beam_scores.clone()})
    for b in range(batch_size):                                                   # finalize :705-725
        for beam_id in range(num_beams):
            bbi = b * num_beams + beam_id
            hyps[b].add(input_ids[bbi].tolist(), beam_scores[bbi].item(), float("nan"))
    out = []
    for h in hyps:                                                                # :555
        out.append([(s * (len(t) ** length_penalty), t, c) for (s, t, c) in h.beams if s > NEG_INF])
    return out


class HFBartStepper:
    """Drives transformers' BartForConditionalGeneration one decoding step at a time in eager fp32
    (what the reference does through model(**model_inputs), :231-238).  Re-runs the decoder over the
    whole prefix each step (t <= 15), which is numerically the cached step and makes
    the beam reorder trivial."""

    def __init__(self, model, input_ids, attention_mask, num_beams):
        self.model = model
        with torch.inference_mode():
            enc = model.get_encoder()(input_ids=input_ids, attention_mask=attention_mask)  # :481-483
        self.enc = enc.last_hidden_state.repeat_interleave(num_beams, dim=0)               # :517-521
        self.mask = attention_mask.repeat_interleave(num_beams, dim=0)

    def __call__(self, decoder_input_ids):
        from transformers.modeling_outputs import BaseModelOutput
        with torch.inference_mode():
            out = self.model(encoder_outputs=BaseModelOutput(last_hidden_state=self.enc), attention_mask=self.mask,
                             decoder_input_ids=decoder_input_ids.to(self.enc.device), use_cache=False)
        return out.logits[:, -1, :].float().cpu()


class HFBartCachedStepper:
    """Same contract as HFBartStepper, but with the decoder KV cache the reference actually runs with
    (`use_cache=True`, seal/beam_search.py:483; `_reorder_cache` after every step, :331-332): each call feeds only the
    newest token and `reorder(beam_idx)` permutes the cache.  This is the honest CPU / eager-GPU baseline -- the
    re-forwarding stepper does O(t^2) decoder work the reference does not do.  Checked against HFBartStepper in
    tests/test_oracle.py (same hypotheses, |dscore| < 1e-5)."""

    def __init__(self, model, input_ids, attention_mask, num_beams):
        self.model = model
        with torch.inference_mode():
            enc = model.get_encoder()(input_ids=input_ids, attention_mask=attention_mask)
        self.enc = enc.last_hidden_state.repeat_interleave(num_beams, dim=0)
        self.mask = attention_mask.repeat_interleave(num_beams, dim=0)
        self.cache = None

    def __call__(self, decoder_input_ids):
        from transformers.modeling_outputs import BaseModelOutput
        with torch.inference_mode():
            dec = decoder_input_ids.to(self.enc.device)
            if self.cache is not None:
                dec = dec[:, -1:]
            out = self.model(encoder_outputs=BaseModelOutput(last_hidden_state=self.enc), attention_mask=self.mask,
                             decoder_input_ids=dec, past_key_values=self.cache, use_cache=True)
            self.cache = out.past_key_values
        return out.logits[:, -1, :].float().cpu()

    def reorder(self, beam_idx):
        with torch.inference_mode():
            self.cache.reorder_cache(beam_idx.to(self.enc.device))


def make_bart(seed=0, device="cpu", layers=None, vocab=None, d_model=None):
    """BartConfig() defaults == facebook/bart-large; seeded random init; the three
    -inf bias entries SEAL sets at load time (seal/retrieval.py:584-588)."""
    from transformers import BartConfig, BartForConditionalGeneration
    kw = {}
    if layers is not None:
        kw.update(encoder_layers=layers, decoder_layers=layers)
    if vocab is not None:
        kw.update(vocab_size=vocab)
    if d_model is not None:
        kw.update(d_model=d_model, encoder_ffn_dim=4 * d_model, decoder_ffn_dim=4 * d_model,
                  encoder_attention_heads=d_model // 64, decoder_attention_heads=d_model // 64)
    cfg = BartConfig(**kw)
    cfg.forced_bos_token_id = None                                                # retrieval.py:566,580
    torch.manual_seed(seed)
    model = BartForConditionalGeneration(cfg).eval().float()
    V = cfg.vocab_size
    with torch.no_grad():
        model.final_logits_bias[0, cfg.pad_token_id] = NEG_INF                    # retrieval.py:586-588
        model.final_logits_bias[0, cfg.bos_token_id] = NEG_INF
        model.final_logits_bias[0, V - 1] = NEG_INF                               # <mask> is the last id
    return model.to(device)


def fm_index_generate_oracle(model, index, input_ids, attention_mask, min_length=3, max_length=25,
                             length_penalty=1.0, num_beams=3, eos_token_id=None, force_decoding_from=None,
                             always_allow_eos=False, disable_fm_index=False, stop_at_count=0,
                             processors=("min_length", "forced_bos", "forced_eos", "inf_nan"), trace=None,
                             use_cache=False, keep_history=True, transformers_output=False, **kw):
    """seal/beam_search.py:391-557 (keep_history=True path) on an HF BART model.  use_cache=True drives the decoder
    with its KV cache like the reference does (baseline timing); the default re-forwards the prefix (simplest exact form)."""
    cfg = model.config
    stepper = (HFBartCachedStepper if use_cache else HFBartStepper)(model, input_ids, attention_mask, num_beams)
    forced_bos = kw.pop("forced_bos_token_id", cfg.forced_bos_token_id)           # :415-418
    return constrained_beam_search_oracle(
        stepper, input_ids.shape[0], index, num_beams, min_length, max_length, length_penalty,
        eos_token_id=eos_token_id if eos_token_id is not None else cfg.eos_token_id,
        pad_token_id=cfg.pad_token_id, decoder_start_token_id=cfg.decoder_start_token_id,
        model_eos_token_id=cfg.eos_token_id, forced_eos_token_id=cfg.forced_eos_token_id,
        forced_bos_token_id=forced_bos, force_decoding_from=force_decoding_from, stop_at_count=stop_at_count,
        always_allow_eos=always_allow_eos, disable_fm_index=disable_fm_index, processors=processors, trace=trace,
        reorder=stepper.reorder if use_cache else None, keep_history=keep_history, transformers_output=transformers_output)
