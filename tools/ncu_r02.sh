#!/bin/bash
# ncu captures of the round-2 build (run under gpurun; outputs in gpurun_out/, summaries exported to profiles/ afterwards)
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# 1. launch lists of the timed steps only (bench.py brackets them with cudaProfilerStart/Stop when SEAL_PROFILE_RANGE is set)
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_q1000.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-big-index > $O/r02_launches_q1000.out 2>&1
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_q20.csv \
    python bench.py --steps 1 --warmup 3 --queries 20 --no-cpu-baseline --no-big-index > $O/r02_launches_q20.out 2>&1
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_q1000_freq.csv \
    python bench.py --steps 1 --warmup 1 --regime freq --no-cpu-baseline --no-big-index > $O/r02_launches_q1000_freq.out 2>&1
# 2. full captures of single launches
$NCU --set full --import-source on -k regex:lf_step_kernel --launch-skip 3 -c 1 -o $O/r02_ncu_lf_1e9 python tools/lf_ncu_probe.py 1000000000 > $O/r02_ncu_lf_1e9.out 2>&1
$NCU --set full --import-source on -k regex:lf_step_kernel --launch-skip 3 -c 1 -o $O/r02_ncu_lf_10m python tools/lf_ncu_probe.py 10000000 > $O/r02_ncu_lf_10m.out 2>&1
$NCU --set full --import-source on -k regex:expand_rows_wide_kernel -c 1 -o $O/r02_ncu_expand_wide python tools/fm_microbench.py > $O/r02_ncu_expand_wide.out 2>&1
$NCU --set full --import-source on -k regex:expand_rows_kernel -c 1 -o $O/r02_ncu_expand_narrow python tools/fm_microbench.py > $O/r02_ncu_expand_narrow.out 2>&1
$NCU --set full --import-source on -k regex:umma_gemm_f16x3_2cta -c 1 -o $O/r02_ncu_2cta_fc1 python tools/gemm_probe.py 5 15000 4096 1024 1 > $O/r02_ncu_2cta_fc1.out 2>&1
$NCU --set full --import-source on -k regex:umma_gemm_f16x3_skinny -c 1 -o $O/r02_ncu_skinny_qkv python tools/gemm_probe.py 5 300 3072 1024 1 > $O/r02_ncu_skinny_qkv.out 2>&1
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --set full --import-source on -k regex:topk_rows_kernel --launch-skip 4 -c 1 -o $O/r02_ncu_topk_rows \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-big-index > $O/r02_ncu_topk_rows.out 2>&1
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --set full --import-source on -k regex:dec_self_attn_kernel --launch-skip 60 -c 1 -o $O/r02_ncu_self_attn \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-big-index > $O/r02_ncu_self_attn.out 2>&1
ls -la $O/*.ncu-rep
