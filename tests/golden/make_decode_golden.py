"""Pins the decode oracle (oracle/decode_oracle.py) against the REFERENCE'S OWN decode code and stores
the reference's outputs as fixtures (tests/golden/decode_golden.json).  Run in the build container only:

    python tests/golden/make_decode_golden.py

/root/reference/seal/beam_search.py is imported UNMODIFIED: fm_index_generate, constrained_beam_search,
IndexBasedLogitsProcessor, BeamSearchScorerWithMemory and BeamHypothesesWithMemory all run as shipped.
What has to be supplied, because transformers 4.13 (requirements.txt:6) is not installed here:

  * stub modules for names that no longer exist (`transformers.generation_utils`, `BeamScorer`, ...) --
    only imported for type annotations / unused branches on this path;
  * `Bart413Adapter`: the HF-4.13 GenerationMixin private helpers fm_index_generate calls
    (beam_search.py:430-445,473-478,481-489,517-521) re-implemented on transformers 5.5's BART with the
    4.13 behaviour SURVEY.md §H3 describes (processor list MinLength -> ForcedBOS -> ForcedEOS ->
    InfNanRemove built from the installed transformers' own processor classes; MaxLengthCriteria;
    full re-forward instead of a KV cache, so `past` stays None).  These are the only assumptions left;
    everything the reference itself implements is pinned by its own code.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.fm_oracle import OracleIndex  # noqa: E402
from oracle.decode_oracle import make_bart, fm_index_generate_oracle  # noqa: E402
from seal_b200.synthetic import make_corpus  # noqa: E402


def load_reference_beam_search():
    import transformers
    from transformers import TopKLogitsWarper
    for name in ("BeamScorer", "BeamSearchScorer", "HammingDiversityLogitsProcessor"):
        if not hasattr(transformers, name):
            setattr(transformers, name, type(name, (), {}))
    gu = types.ModuleType("transformers.generation_utils")
    for name in ("BeamSearchOutput", "BeamSearchEncoderDecoderOutput", "BeamSearchDecoderOnlyOutput"):
        setattr(gu, name, type(name, (), {}))
    gu.validate_stopping_criteria = lambda sc, max_length: sc
    glp = types.ModuleType("transformers.generation_logits_process"); glp.TopKLogitsWarper = TopKLogitsWarper
    mi = types.ModuleType("more_itertools"); mi.chunked = lambda it, n: (it[i:i + n] for i in range(0, len(it), n))
    seal = types.ModuleType("seal"); seal_index = types.ModuleType("seal.index"); seal_index.FMIndex = OracleIndex
    seal.index = seal_index; seal.FMIndex = OracleIndex
    for k, v in {"transformers.generation_utils": gu, "transformers.generation_logits_process": glp,
                 "more_itertools": mi, "seal": seal, "seal.index": seal_index}.items():
        sys.modules.setdefault(k, v)
    spec = importlib.util.spec_from_file_location("ref_beam_search", "/root/reference/seal/beam_search.py")
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


class _MaxLength:
    """StoppingCriteriaList([MaxLengthCriteria(max_length)]) with the 4.13 return type (one bool)."""
    def __init__(self, max_length): self.max_length = max_length
    def __len__(self): return 1
    def __call__(self, input_ids, scores, **kw): return input_ids.shape[-1] >= self.max_length


class _InfNanRemove413:
    """InfNanRemoveLogitsProcessor as of transformers 4.13: NaN -> 0, +inf -> finfo.max.  (The installed 5.5
    class additionally maps -inf -> finfo.min, a later change that would let masked candidates survive
    the reference's `h[0] > -inf` filter, beam_search.py:555.)"""
    def __call__(self, input_ids, scores):
        scores[scores != scores] = 0.0
        scores[scores == float("inf")] = torch.finfo(scores.dtype).max
        return scores


class Bart413Adapter:
    def __init__(self, model):
        self.m = model
        c = model.config
        self.config = types.SimpleNamespace(
            pad_token_id=c.pad_token_id, eos_token_id=c.eos_token_id, bos_token_id=c.bos_token_id,
            decoder_start_token_id=c.decoder_start_token_id, forced_bos_token_id=c.forced_bos_token_id,
            forced_eos_token_id=c.forced_eos_token_id, is_encoder_decoder=True, output_scores=False,
            output_attentions=False, output_hidden_states=False, return_dict_in_generate=False, vocab_size=c.vocab_size)

    # generation_utils.py (4.13) _get_logits_processor: None arguments fall back to model.config
    def _get_logits_processor(self, min_length=None, max_length=None, eos_token_id=None, forced_bos_token_id=None,
                              forced_eos_token_id=None, remove_invalid_values=None, **unused):
        from transformers import (LogitsProcessorList, MinLengthLogitsProcessor, ForcedBOSTokenLogitsProcessor,
                                  ForcedEOSTokenLogitsProcessor)
        eos = eos_token_id if eos_token_id is not None else self.config.eos_token_id
        fbos = forced_bos_token_id if forced_bos_token_id is not None else self.config.forced_bos_token_id
        feos = forced_eos_token_id if forced_eos_token_id is not None else self.config.forced_eos_token_id
        procs = LogitsProcessorList()
        if min_length is not None and eos is not None and min_length > -1:
            procs.append(MinLengthLogitsProcessor(min_length, eos, device="cpu"))
        if fbos is not None:
            procs.append(ForcedBOSTokenLogitsProcessor(fbos))
        if feos is not None:
            procs.append(ForcedEOSTokenLogitsProcessor(max_length, feos, device="cpu"))
        if remove_invalid_values:
            procs.append(_InfNanRemove413())
        return procs

    def _get_stopping_criteria(self, max_length=None, max_time=None):
        return _MaxLength(max_length)

    def _prepare_encoder_decoder_kwargs_for_generation(self, input_ids, model_kwargs):
        enc = self.m.get_encoder()(input_ids=input_ids, attention_mask=model_kwargs["attention_mask"], return_dict=True)
        model_kwargs["encoder_outputs"] = enc
        return model_kwargs

    def _prepare_decoder_input_ids_for_generation(self, batch_size, decoder_start_token_id=None, bos_token_id=None):
        return torch.full((batch_size, 1), decoder_start_token_id, dtype=torch.long)

    def _expand_inputs_for_generation(self, input_ids, expand_size=1, is_encoder_decoder=False, attention_mask=None,
                                      encoder_outputs=None, **model_kwargs):
        idx = torch.arange(input_ids.shape[0]).view(-1, 1).repeat(1, expand_size).view(-1)
        model_kwargs["attention_mask"] = attention_mask.index_select(0, idx)
        encoder_outputs["last_hidden_state"] = encoder_outputs.last_hidden_state.index_select(0, idx)
        model_kwargs["encoder_outputs"] = encoder_outputs
        return input_ids.index_select(0, idx), model_kwargs

    def prepare_inputs_for_generation(self, decoder_input_ids, attention_mask=None, encoder_outputs=None, **kw):
        return dict(decoder_input_ids=decoder_input_ids, attention_mask=attention_mask, encoder_outputs=encoder_outputs)

    def __call__(self, decoder_input_ids=None, attention_mask=None, encoder_outputs=None, **kw):
        return self.m(decoder_input_ids=decoder_input_ids, attention_mask=attention_mask, encoder_outputs=encoder_outputs,
                      use_cache=False, return_dict=True)

    def adjust_logits_during_generation(self, logits, cur_len=None):      # BART, 4.13: identity
        return logits

    def _update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False):
        model_kwargs["past"] = None                                       # full re-forward: nothing to reorder
        return model_kwargs


def make_inputs(rng, Q, S, vocab):
    ids = torch.tensor(rng.integers(4, vocab, size=(Q, S)), dtype=torch.long)
    am = torch.ones_like(ids)
    ids[:, 0] = 0
    for q in range(Q):
        l = int(rng.integers(max(3, S // 2), S + 1))
        ids[q, l - 1] = 2
        ids[q, l:] = 1
        am[q, l:] = 0
    return ids, am


CORPUS = dict(n_docs=300, doc_len=30, n_phrases=600, seed=3, vocab=2000)
MODEL = dict(seed=0, layers=2, vocab=2000, d_model=128)
CASES = [
    dict(num_beams=5, min_length=8, max_length=8, length_penalty=0.0),
    dict(num_beams=3, min_length=2, max_length=6, length_penalty=1.0),
    dict(num_beams=4, min_length=0, max_length=7, length_penalty=0.0, always_allow_eos=True),
    dict(num_beams=4, min_length=0, max_length=7, length_penalty=0.0, stop_at_count=3),
    dict(num_beams=5, min_length=3, max_length=7, length_penalty=0.0, force_decoding_from=[996, 523]),
    dict(num_beams=3, min_length=0, max_length=6, length_penalty=0.0, forced_bos_token_id=0),
    dict(num_beams=4, min_length=0, max_length=6, length_penalty=0.0, disable_fm_index=True),
    dict(num_beams=15, min_length=10, max_length=10, length_penalty=0.0),
    dict(num_beams=5, min_length=0, max_length=9, length_penalty=0.0, eos_token_id=777, force_decoding_from=[2]),
]


def main():
    ref = load_reference_beam_search()
    docs = make_corpus(**CORPUS)
    ora = OracleIndex([d.tolist() for d in docs], backend="ref")
    model = make_bart(**MODEL)
    adapter = Bart413Adapter(model)
    out = {"corpus": CORPUS, "model": MODEL, "cases": []}
    worst_all = 0.0
    for ci, kw in enumerate(CASES):
        rng = np.random.default_rng(100 + ci)
        ids, am = make_inputs(rng, Q=4, S=12, vocab=CORPUS["vocab"])
        got_ref = ref.fm_index_generate(adapter, ora, ids, am, keep_history=True, **kw)
        got_ora = fm_index_generate_oracle(model, ora, ids, am, **kw)
        # the oracle keeps -inf-scored records out as the reference does (:555); compare complete lists, in order
        worst = 0.0
        for q, (a, b) in enumerate(zip(got_ref, got_ora)):
            ta = [tuple(t) for _, t in a]; tb = [tuple(t) for _, t, _ in b]
            assert ta == tb, f"case {ci} query {q}: hypothesis lists differ\nref  {ta[:6]}\nours {tb[:6]}"
            for (sa, _), (sb, _, _) in zip(a, b):
                worst = max(worst, abs(sa - sb))
        worst_all = max(worst_all, worst)
        print(f"case {ci} {kw}: {sum(len(a) for a in got_ref)} hypotheses identical in order, worst |dscore| {worst:.2e}")
        out["cases"].append({"kw": kw, "seed": 100 + ci, "input_ids": ids.tolist(), "attention_mask": am.tolist(),
                             "hyps": [[[float(s), [int(x) for x in t]] for s, t in a] for a in got_ref]})
    assert worst_all < 1e-5
    with open(os.path.join(HERE, "decode_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote decode_golden.json", os.path.getsize(os.path.join(HERE, "decode_golden.json")), "bytes")


def fuzz(n_cases):
    """Randomised cross-check (nothing stored): the reference's fm_index_generate vs the oracle restatement on random
    parameter combinations -- same hypotheses in the same order, |dscore| < 1e-5, or the same exception type."""
    ref = load_reference_beam_search()
    docs = make_corpus(**CORPUS)
    ora = OracleIndex([d.tolist() for d in docs], backend="ref")
    model = make_bart(**MODEL)
    adapter = Bart413Adapter(model)
    rng = np.random.default_rng(31337)
    bad = raised = 0
    for case in range(n_cases):
        max_length = int(rng.integers(3, 11))
        kw = dict(num_beams=int(rng.integers(1, 9)), max_length=max_length, min_length=int(rng.integers(0, max_length + 1)),
                  length_penalty=float(rng.choice([0.0, 0.5, 1.0])))
        if rng.random() < 0.3: kw["always_allow_eos"] = True
        if rng.random() < 0.3: kw["stop_at_count"] = int(rng.choice([1, 2, 5]))
        if rng.random() < 0.25:
            d = int(rng.integers(0, docs.shape[0])); a = int(rng.integers(0, docs.shape[1] - 3))
            kw["force_decoding_from"] = [int(t) for t in docs[d, a:a + int(rng.integers(1, 3))]]
        if rng.random() < 0.2: kw["forced_bos_token_id"] = 0
        if rng.random() < 0.15: kw["disable_fm_index"] = True
        if rng.random() < 0.2: kw["eos_token_id"] = int(rng.integers(4, CORPUS["vocab"]))
        ids, am = make_inputs(rng, Q=int(rng.integers(1, 4)), S=int(rng.integers(4, 13)), vocab=CORPUS["vocab"])
        try:
            a = ref.fm_index_generate(adapter, ora, ids, am, keep_history=True, **kw)
        except Exception as e:
            raised += 1
            try:
                fm_index_generate_oracle(model, ora, ids, am, **kw)
                print("case", case, kw, "reference raised", type(e).__name__, e, "but the oracle did not"); bad += 1
            except Exception as e2:
                if type(e2) is not type(e):
                    print("case", case, kw, "different exceptions", type(e).__name__, type(e2).__name__); bad += 1
            continue
        b = fm_index_generate_oracle(model, ora, ids, am, **kw)
        ok = len(a) == len(b)
        for qa, qb in zip(a, b):
            ok = ok and [tuple(t) for _, t in qa] == [tuple(t) for _, t, _ in qb]
            ok = ok and all(abs(x[0] - y[0]) < 1e-5 for x, y in zip(qa, qb))
        if not ok:
            bad += 1
            print("case", case, "MISMATCH", kw)
    print(f"fuzz: {n_cases} cases ({raised} where both raise), {bad} mismatches")
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--fuzz":
        sys.exit(1 if fuzz(int(sys.argv[2])) else 0)
    main()
