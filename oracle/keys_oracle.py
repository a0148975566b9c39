"""TEST INFRASTRUCTURE ONLY.  CPU/torch restatement of seal/keys.py:64-141 (rescore_keys) and
:145-176 (compute_unigram_scores) for the installed transformers (the reference file imports
`more_itertools`/`seal` and uses HF-4.13 private helpers, so it is restated, line references kept)."""
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
