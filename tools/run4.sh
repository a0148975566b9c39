#!/bin/bash
# GPU run 4: split in two stages so that a kernel fault in the tests cannot spoil the measurements
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/r02_t4.log; tail -4 $O/r02_t4.log
python bench.py --steps 30 --warmup 5 --queries 20 --no-cpu-baseline --no-big-index > $O/r02_bench4_q20.json 2> $O/r02_bench4_q20.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-big-index > $O/r02_bench4_q1000.json 2> $O/r02_bench4_q1000.err
SEALB200_SELF_ATTN_QUERY=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-big-index > $O/r02_bench4_q1000_oldattn.json 2> $O/r02_bench4_q1000_oldattn.err
python bench.py --steps 3 --warmup 3 --regime freq --no-cpu-baseline --no-big-index > $O/r02_bench4_freq.json 2> $O/r02_bench4_freq.err
python tools/fm_microbench.py > $O/r02_fm_microbench4.log 2>&1; cp $O/fm_microbench.json $O/r02_fm_microbench4.json
python - <<'PY'
import json
for n in ("q20","q1000","q1000_oldattn","freq"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench4_{n}.json").read().splitlines()[-1]); print(n, round(d["ms_per_step"],2), round(d["value"],1), d["phases_us_last_step"])
    except Exception as e: print(n, "ERR", e)
m=json.load(open("gpurun_out/r02_fm_microbench4.json")); print({k:round(v["expand_us"]) for k,v in m["walk_R15000"].items()})
PY
python bench.py --steps 5 --warmup 3 > $O/r02_bench4_full.json 2> $O/r02_bench4_full.err; tail -c 300 $O/r02_bench4_full.err
