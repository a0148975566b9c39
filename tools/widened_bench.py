"""Measurement for the widened rows on the benchmark corpus: after one constrained
generate pass for Q queries (beam 15, body defaults) the keys go through the reference's post-processing
(retrieval.py:86-91), then
  * rescore_keys + compute_unigram_scores  (seal/keys.py:64-176)  : product on the GPU vs the torch restatement
    on host cores (bounded sample of queries);
  * aggregate_evidence                      (seal/keys.py:178-497): product (three batched index launches per
    query) vs the sequential restatement on the compiled reference FM-index (one call per key / unigram / row /
    document, like the reference).
Prints one JSON object.  Usage: python tools/widened_bench.py [Q=20] [cpu_queries=3]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from bench import make_model, build_inputs
    from seal_b200.beam_search import fm_index_generate
    from seal_b200.cpp_modules.fm_index import FMIndex as RawFM
    from seal_b200.index import FMIndex
    from seal_b200.keys import rescore_keys, compute_unigram_scores, aggregate_evidence
    from seal_b200.synthetic import corpus_symbols
    from oracle.fm_oracle import OracleIndex, RefFM, PortFM, ref_available
    from oracle.keys_oracle import rescore_keys_oracle, compute_unigram_scores_oracle, aggregate_evidence_oracle

    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    QC = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    docs, ids_np, mask_np = build_inputs(Q, seed=4321)
    t0 = time.perf_counter()
    index = FMIndex(); RawFM.initialize(index, corpus_symbols(docs))
    index.beginnings = list(range(0, docs.size + 1, docs.shape[1])); index._sync_beginnings(); index.to_device(0)
    index.occurring_distinct, index.occurring_counts = index.get_distinct_count(0, len(index))
    t_build = time.perf_counter() - t0
    model = make_model()
    ids = torch.from_numpy(ids_np); mask = torch.from_numpy(mask_np)
    gen = lambda: fm_index_generate(model, index, ids, mask, min_length=10, max_length=10, length_penalty=0.0, num_beams=15,
                                    keep_history=True)
    found = gen()
    torch.cuda.synchronize(); t0 = time.perf_counter(); found = gen(); torch.cuda.synchronize(); t_gen = time.perf_counter() - t0

    strip = {0, 2}                                                   # retrieval.py:86-91 with bos/eos as strip tokens
    cnt = {}
    keys = []
    for fk in found:
        fk = [(s, k[1:] if k and k[0] in strip else k) for s, k in fk if k]
        fk = [(s, k[1:] if k and k[0] in strip else k) for s, k in fk if k]
        fk = [(s, k[:-1] if k and k[-1] in strip else k) for s, k in fk if k]
        flat = [k for _, k in fk if k]
        lo, hi = index.get_range_batch(flat) if flat else ([], [])
        ok = {tuple(k) for k, a, b in zip(flat, lo, hi) if int(b) > int(a)}
        seen = set(); out = []
        for s, k in fk:
            if k and tuple(k) in ok and tuple(k) not in seen:
                seen.add(tuple(k)); out.append((s, k))
        keys.append(out)
    inputs = [ids_np[q, :int(mask_np[q].sum())].tolist() for q in range(Q)]
    res = {"queries": Q, "index_build_s": round(t_build, 2), "generate_s": round(t_gen, 4),
           "keys_per_query": float(np.mean([len(k) for k in keys])), "host_cores": os.cpu_count()}

    # ---- rank 1: teacher-forced scoring ----------------------------------------------------------------
    rescore_keys(model, inputs, keys)                                # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    scored = rescore_keys(model, inputs, keys)
    uni = compute_unigram_scores(model, inputs)
    torch.cuda.synchronize(); t_tf = time.perf_counter() - t0
    nkeys = sum(len(k) for k in keys)
    t0 = time.perf_counter()
    exp = rescore_keys_oracle(model, inputs[:QC], keys[:QC])
    exp_uni = compute_unigram_scores_oracle(model, inputs[:QC])
    t_tf_cpu = time.perf_counter() - t0
    worst = max([abs(a[0] - b[0]) for qa, qb in zip(scored[:QC], exp) for a, b in zip(qa, qb)] + [0.0])
    res["rescore"] = {"gpu_s": round(t_tf, 4), "keys": nkeys, "gpu_keys_per_s": round(nkeys / t_tf, 1),
                      "cpu_sample_queries": QC, "cpu_s": round(t_tf_cpu, 3),
                      "cpu_keys_per_s": round(sum(len(k) for k in keys[:QC]) / t_tf_cpu, 1), "worst_abs_dscore_vs_oracle": worst,
                      "cpu_kind": "torch restatement of keys.py:64-176 on host cores (eager fp32)"}

    # ---- rank 2: evidence aggregation ------------------------------------------------------------------
    keys2 = [[(k, s) for s, k in q] for q in scored]                 # (ngram, score) pairs as retrieval.py passes them
    aggregate_evidence(keys2[0], unigram_scores=list(uni[0]), index=index)            # warm-up
    t0 = time.perf_counter()
    outs = [aggregate_evidence(keys2[q], unigram_scores=list(uni[q]), index=index) for q in range(Q)]
    t_ev = time.perf_counter() - t0
    t0 = time.perf_counter()
    raw = RefFM(corpus_symbols(docs)) if ref_available() else PortFM(corpus_symbols(docs))
    ora = OracleIndex(_raw=raw); ora.beginnings = index.beginnings
    t_ora_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    exps = [aggregate_evidence_oracle(keys2[q], unigram_scores=list(uni[q]), index=ora) for q in range(QC)]
    t_ev_cpu = time.perf_counter() - t0
    same = all(list(a[0].keys()) == list(b[0].keys()) and [v[0] for v in a[0].values()] == [v[0] for v in b[0].values()]
               and a[1] == b[1] for a, b in zip(outs[:QC], exps))
    res["aggregate_evidence"] = {"gpu_batched_s_per_query": round(t_ev / Q, 4), "queries": Q,
                                 "docs_scored_per_query": float(np.mean([len(o[0]) for o in outs])),
                                 "sequential_s_per_query": round(t_ev_cpu / QC, 3), "sequential_sample_queries": QC,
                                 "identical_to_sequential": bool(same), "reference_index_build_s": round(t_ora_build, 1),
                                 "sequential_kind": ("compiled reference FM-index (oracle/_ref) behind " if ref_available() else "C port behind ")
                                 + "the restated keys.py:178-497, one index call per key / unigram / row / document"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
