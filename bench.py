#!/usr/bin/env python
"""Headline benchmark: SEAL constrained beam-search decode, queries/sec at beam 15 on a synthetic
10 M-token FM-index with BART-large (BASELINE.json metric / configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on host cores

One "step" = one full pass of the hot path (encoder, 9 constrained decode steps, hypothesis
records) over one batch of `--queries` synthetic queries.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec at beam=15, 1k queries, 10M-token index; rank-kernel HBM GB/s"
BEAM, MIN_LEN, MAX_LEN, LP = 15, 10, 10, 0.0          # SEALSearcher body defaults (retrieval.py:70-83)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the fc1-shaped GEMM (M=15000, N=4096, K=1024) from
# `ncu --set full` (profiles/r01_ncu_f16_fc1_v2_raw.csv: 80.0 MB read + 210.8 MB written; r01_ncu_umma_fc1_raw.csv for
# the TF32 kernel); algorithmic bytes of that launch: A halves 61 MB + W halves 17 MB + C halves 246 MB = 324 MB, of
# which the activations/weights mostly hit L2 (they were just written by the producer kernel).
TRAFFIC_PER_LAUNCH = {2: 438.3e6, 3: 290.8e6, 4: 290.8e6, 5: 292.1e6}   # dram read+write of the fc1-shaped launch (ncu --set full, profiles/)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(n_queries, seed):
    from seal_b200.synthetic import make_corpus, make_queries
    docs = make_corpus()                                     # 100 000 docs x 100 tokens, seed 1234
    ids, mask = make_queries(n_queries, seed=seed)
    return docs, ids, mask


def make_model():
    import torch
    from transformers import BartConfig, BartForConditionalGeneration
    cfg = BartConfig()                                       # defaults == facebook/bart-large
    cfg.forced_bos_token_id = None                           # seal/retrieval.py:566,580
    torch.manual_seed(0)
    model = BartForConditionalGeneration(cfg).eval().float()
    with torch.no_grad():
        for t in (cfg.pad_token_id, cfg.bos_token_id, cfg.vocab_size - 1):
            model.final_logits_bias[0, t] = float("-inf")    # seal/retrieval.py:584-588
    return model


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ctypes as C
    from seal_b200._lib import lib, check
    from seal_b200.beam_search import SealBartEngine, _make_params, _occurring_mask, generate_records
    from seal_b200.index import FMIndex
    from seal_b200.synthetic import corpus_symbols

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    Q = args.queries
    docs, ids_np, mask_np = build_inputs(Q, seed=4321 + rank)          # every rank: its own 1k queries
    # build straight from the symbol stream (same result as FMIndex.initialize, without 100k Python lists)
    from seal_b200.cpp_modules.fm_index import FMIndex as RawFM
    index = FMIndex()
    RawFM.initialize(index, corpus_symbols(docs))
    index.beginnings = list(range(0, docs.size + 1, docs.shape[1]))
    index._sync_beginnings()
    index.to_device(local)
    index.occurring_distinct, index.occurring_counts = index.get_distinct_count(0, len(index))
    model = make_model()
    eng = SealBartEngine.from_hf(model, device=local, gemm_mode=args.gemm_mode)
    cfg = model.config
    del model
    p = _make_params(cfg, BEAM, MIN_LEN, MAX_LEN, LP, cfg.eos_token_id, None, False, False, 0, None)
    H = int(lib.sealdec_hyps_per_query(C.byref(p))); T = MAX_LEN
    S = ids_np.shape[1]
    occ_np = _occurring_mask(index, cfg.vocab_size)
    occ = torch.from_numpy(occ_np.view(np.int32)).to(dev)
    ids = torch.from_numpy(ids_np).to(dev); mask = torch.from_numpy(mask_np).to(dev)
    o_score = torch.empty((Q, H), dtype=torch.float32, device=dev); o_len = torch.empty((Q, H), dtype=torch.int32, device=dev)
    o_tok = torch.empty((Q, H, T), dtype=torch.int32, device=dev); o_valid = torch.empty((Q, H), dtype=torch.uint8, device=dev)
    o_lo = torch.empty((Q, H), dtype=torch.int64, device=dev); o_hi = torch.empty((Q, H), dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    rec_bytes = Q * H * (4 + 4 + 4 * T + 1 + 16)
    gathered = [torch.empty_like(o_score) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step_device(gather=True):
        st = torch.cuda.current_stream().cuda_stream
        check(lib.sealdec_generate_d(eng._h, index._dev(), occ.data_ptr(), C.byref(p), ids.data_ptr(), mask.data_ptr(),
                                     Q, S, st, o_score.data_ptr(), o_len.data_ptr(), o_tok.data_ptr(), o_valid.data_ptr(),
                                     o_lo.data_ptr(), o_hi.data_ptr(), err.data_ptr()))
        if world > 1 and gather:                        # the single collective: result records to rank 0
            dist.gather(o_score, gathered, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    prof_range = bool(os.environ.get("SEAL_PROFILE_RANGE"))   # ncu --profile-from-start off: capture the timed steps only
    if prof_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    if prof_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = eng.last_launch_count() * args.steps
    phases = eng.last_phase_us()
    assert int(err.item()) == 0
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * Q * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (H2D inputs + D2H records inside) ----------
    pin_ids = torch.from_numpy(ids_np).pin_memory(); pin_mask = torch.from_numpy(mask_np).pin_memory()
    for _ in range(min(args.warmup, 2)):
        generate_records(eng, index, pin_ids.numpy(), pin_mask.numpy(), MIN_LEN, MAX_LEN, LP, BEAM, forced_bos_token_id=None)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        rec = generate_records(eng, index, pin_ids.numpy(), pin_mask.numpy(), MIN_LEN, MAX_LEN, LP, BEAM, forced_bos_token_id=None)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    e2e = {"value": world * Q * e2e_steps / e2e_s, "unit": "queries/s",
           "h2d_bytes_per_step": int(ids_np.nbytes + mask_np.nbytes + occ_np.nbytes),
           "d2h_bytes_per_step": int(rec_bytes + 4), "api": "sealdec_generate (host buffers)"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    hbm, tf_burst, tf_sus, which = peaks()
    # ---- roofline of the dominant kernel (the decoder/encoder/lm_head GEMM), measured live: one extra
    # pass with every GEMM launch bracketed by CUDA events on its stream --------------------------------
    eng.profile_gemm(True)
    step_device(gather=False); torch.cuda.synchronize()
    prof = eng.profile_gemm(False)
    gemm_s = prof["total_us"] * 1e-6
    passes = {0: 1, 1: 3, 2: 3, 3: 3, 4: 3, 5: 3}[args.gemm_mode]
    ach = prof["flops"] / gemm_s / 1e12
    roof = {"bound": "tensor", "kernel": {0: "sgemm_tn_kernel", 1: "umma_gemm_tf32x3_kernel", 2: "umma_gemm_tf32x3_persistent_kernel",
                                          3: "umma_gemm_f16x3_persistent_kernel", 4: "umma_gemm_f16x3_persistent_kernel<ROWB=64>", 5: "umma_gemm_f16x3_2cta_kernel"}[args.gemm_mode],
            "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
            "traffic": TRAFFIC_PER_LAUNCH.get(args.gemm_mode),
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({which}; the kernel runs inside a long step)",
            "avg_launch_us": prof["total_us"] / max(prof["launches"], 1), "launches_per_step": prof["launches"],
            "share_of_step": gemm_s / (ms / args.steps * 1e-3),
            "tensor_pipe_TFLOPs": ach * passes, "tensor_pipe_frac": ach * passes / tf_sus,
            "note": "achieved = algorithmic 2MNK flops (fp32-equivalent) of all GEMM launches of one step / their summed "
                    "CUDA-event durations; the kernel issues 3 half-precision tensor-core passes per product "
                    "(error-compensated split, DESIGN.md section 4), so the tensor pipe itself runs at tensor_pipe_TFLOPs; "
                    "traffic = dram bytes read+written per launch of the fc1-shaped GEMM (ncu, profiles/)"}
    # ---- the metric's rank kernel: batched LF-mapping (backward_search_step) on this 10 M-token index,
    # 4 M random (symbol, lo, hi) triples, CUDA events.  The 27 MB index is L2-resident, so this is L2,
    # not HBM, bandwidth; the beyond-L2 figure is in profiles/r01_fm_microbench_v3_bigindex.json.
    Nlf = 1 << 22
    g = torch.Generator(device=dev); g.manual_seed(1)
    sym = torch.randint(14, 50000, (Nlf,), device=dev, generator=g)
    lo_t = torch.randint(0, index.size() // 2, (Nlf,), device=dev, generator=g)
    hi_t = lo_t + torch.randint(1, index.size() // 2, (Nlf,), device=dev, generator=g)
    for _ in range(3):
        index.lf_step_tensors(sym, lo_t, hi_t)
    ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(10):
        index.lf_step_tensors(sym, lo_t, hi_t)
    eb.record(); torch.cuda.synchronize()
    lf_s = ea.elapsed_time(eb) * 1e-3 / 10
    rank_kernel = {"kernel": "lf_step_kernel", "triples": Nlf, "us": lf_s * 1e6, "steps_per_s": Nlf / lf_s,
                   "algorithmic_GBps": Nlf * 48 * 16 / lf_s / 1e9, "hbm_peak_GBps": hbm,
                   "frac_of_hbm_peak": Nlf * 48 * 16 / lf_s / 1e9 / hbm,
                   "note": "48*L B per LF step (SURVEY 8d); index L2-resident at 10 M tokens; "
                           "select+expand phase of the step: %.1f ms of %.1f ms" % (phases["select_expand"] / 1e3, phases["total"] / 1e3)}
    out = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "configs[1]: synthetic 10M-token corpus (100k docs x 100 tok, seed 1234), "
                                  f"{Q} queries/GPU (seed 4321+rank), beam {BEAM}, min=max_length {MAX_LEN}, BART-large "
                                  "random init seed 0, fp32", "queries_per_gpu": Q, "beam": BEAM, "decode_steps": MAX_LEN - 1,
                      "parallelism": f"query-sharded x{world}, index+weights replicated, one NCCL gather",
                      "l2": "per-step working set (KV cache + logits > 10 GB) exceeds L2; no explicit flush",
                      "exact_work_elision": "results identical to the full computation (parity tests): the first decode step runs on one "
                                            "row per query (its beams are identical rows) [SEALB200_COMPACT_FIRST=%s]; the step whose scores "
                                            "ForcedEOS overwrites entirely (the 9th) has no model forward [SEALB200_SKIP_DEAD_STEP=%s]; "
                                            "all 9 select/record steps run; the encoder runs on the real (unpadded) source tokens [SEALB200_PACK_ENCODER=%s]"
                                            % (os.environ.get("SEALB200_COMPACT_FIRST", "1"), os.environ.get("SEALB200_SKIP_DEAD_STEP", "1"),
                                               os.environ.get("SEALB200_PACK_ENCODER", "1"))},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
           "roofline": roof, "rank_kernel": rank_kernel, "phases_us_last_step": phases,
           "cpu_baseline": cpu_baseline_sample(args) if world == 1 else None}   # rank 0 at N = 1 only
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
def reference_setup(n_queries, seed=4321):
    """The reference algorithm on host cores: CPU restatement of seal/beam_search.py (oracle/
    decode_oracle.py) on transformers' eager fp32 BART + the reference FM-index (oracle/_ref, the
    unmodified seal/cpp_modules/fm_index.cpp on sdsl-lite; the C port if _ref was not shipped)."""
    import torch
    from oracle.fm_oracle import OracleIndex, RefFM, PortFM, ref_available
    from seal_b200.synthetic import corpus_symbols
    docs, ids, mask = build_inputs(n_queries, seed)
    fm = RefFM(corpus_symbols(docs)) if ref_available() else PortFM(corpus_symbols(docs))
    idx = OracleIndex(_raw=fm)
    idx.beginnings = list(range(0, docs.size + 1, docs.shape[1]))
    idx.occurring_distinct, idx.occurring_counts = idx.get_distinct_count(0, len(idx))
    model = make_model()
    return idx, model, torch.from_numpy(ids), torch.from_numpy(mask), ("reference" if ref_available() else "port")


def reference_step(idx, model, ids, mask, lo, n):
    from oracle.decode_oracle import fm_index_generate_oracle
    return fm_index_generate_oracle(model, idx, ids[lo:lo + n], mask[lo:lo + n], min_length=MIN_LEN, max_length=MAX_LEN,
                                    length_penalty=LP, num_beams=BEAM)


def cpu_baseline_sample(args):
    if args.no_cpu_baseline:
        return None
    try:
        import torch
        n = args.ref_queries
        idx, model, ids, mask, kind = reference_setup(max(n, 1))
        t0 = time.perf_counter()
        reference_step(idx, model, ids, mask, 0, n)
        dt = time.perf_counter() - t0
        return {"value": n / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": kind,
                "sample": f"{n} of the 1000 queries, full 9-step constrained decode (HF BART eager fp32 on CPU + "
                          f"{'sdsl-lite FM-index (oracle/_ref)' if kind == 'reference' else 'C port of the FM-index'}), {dt:.1f} s"}
    except Exception as ex:  # pragma: no cover
        return {"value": None, "unit": "queries/s", "error": repr(ex)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    n = args.ref_queries
    idx, model, ids, mask, kind = reference_setup(n * (args.steps + args.warmup))
    k = 0
    for _ in range(args.warmup):
        reference_step(idx, model, ids, mask, k, n); k += n
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_step(idx, model, ids, mask, k, n); k += n
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    base = {"value": v, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{n} queries per step (bounded sample of the 1000-query batch), 9 decode steps, beam {BEAM}"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s",
                      "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "configs[1] sample: synthetic 10M-token corpus, beam 15, min=max_length 10, "
                                             "BART-large random init seed 0, fp32, host cores only", "queries_per_step": n},
                      "cpu_baseline": base,
                      "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--ref-queries", type=int, default=2, help="queries per step of the CPU reference sample")
    ap.add_argument("--gemm-mode", type=int, default=int(os.environ.get("SEALB200_GEMM", "5")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
