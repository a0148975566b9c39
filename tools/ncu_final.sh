#!/bin/bash
# launch lists + one full capture of the FINAL build of the round (profiles/r02_f_*)
O=gpurun_out
NCU="ncu --clock-control none"
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02f_launches_q1000.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-big-index > $O/r02f_launches_q1000.out 2>&1
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r02f_launches_q20.csv \
    python bench.py --steps 1 --warmup 3 --queries 20 --no-cpu-baseline --no-big-index > $O/r02f_launches_q20.out 2>&1
SEAL_PROFILE_RANGE=1 $NCU --profile-from-start off --set full --import-source on -k regex:dec_self_attn_query --launch-skip 60 -c 1 -o $O/r02f_ncu_self_attn_query \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-big-index > $O/r02f_ncu_self_attn_query.out 2>&1
ls -la $O/r02f_*
