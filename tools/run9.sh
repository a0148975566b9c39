#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02_t9.log; tail -3 $O/r02_t9.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench9_full.json 2> $O/r02_bench9_full.err; tail -c 300 $O/r02_bench9_full.err
python bench.py --steps 30 --warmup 5 --queries 20 --no-cpu-baseline --no-big-index > $O/r02_bench9_q20.json 2> $O/r02_bench9_q20.err
python tools/baselines.py --out $O/r02_baselines9.json > $O/r02_baselines9.log 2>&1; tail -c 500 $O/r02_baselines9.log
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench9_full.json').read().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check']['ok'], d['cpu_baseline'], d['rank_kernel']['hbm_index'])
d=json.loads(open('gpurun_out/r02_bench9_q20.json').read().splitlines()[-1]); print('q20', d['ms_per_step'], d['e2e']['value'])
"
