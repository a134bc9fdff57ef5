#!/bin/bash
# whole GPU suite + the bench line (N=1) + a few probe points
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/t4_bench.json 2> gpurun_out/t4_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/t4_bench.err
{
for t in 51 251 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 | tail -1; done
for t in 51 251 1001; do I24=1 timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 | tail -1; done
} > gpurun_out/tc2_probe_final.txt 2>&1
cut -c1-200 gpurun_out/tc2_probe_final.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t4_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['value_sustained']['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['tensor'])
for k,v in d['configs'].items(): print(k, v['value'], v['e2e']['value'])
PY
