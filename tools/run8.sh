#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r02_t8.log; tail -3 $O/r02_t8.log
python bench.py --steps 10 --warmup 3 > $O/r02_bench8_full.json 2> $O/r02_bench8_full.err; tail -c 300 $O/r02_bench8_full.err
python bench.py --steps 20 --warmup 5 --impl reference > $O/r02_bench8_reference.json 2> $O/r02_bench8_reference.err; tail -c 300 $O/r02_bench8_reference.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench8_full.json').read().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity_check']['ok'], d['rank_kernel']['hbm_index'])
r=json.loads(open('gpurun_out/r02_bench8_reference.json').read().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
"
free -g | head -2
MEM=$(free -g | awk '/Mem:/ {print $7}')
python tools/big_index_bench.py --out $O/r02_big_index8.json > $O/r02_big_index8.log 2>&1; tail -c 400 $O/r02_big_index8.log
if [ "$MEM" -gt 300 ]; then timeout 900 python tools/big_index_bench.py --tokens 2200000000 --no-decode --out $O/r02_big_index_2p2e9.json > $O/r02_big_index_2p2e9.log 2>&1; tail -c 1200 $O/r02_big_index_2p2e9.log; fi
