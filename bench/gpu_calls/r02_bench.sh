#!/bin/bash
timeout 900 python bench.py > gpurun_out/t5_bench.json 2> gpurun_out/t5_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/t5_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t5_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['value_sustained']['value'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, v['value'], v['e2e']['value'], v['roofline'].get('bound'), v['roofline'].get('frac'), v['stage'])
for r in d['c5_sweep']:
    if r['streams_total'] in (16384,65536): print(r['taps'], r['upsample'], r['streams_total'], r['frames'], r.get('blocks_per_call',1), r['gsamples_s'], r['frac'], r['hbm_frac'], r['kernel'])
PY
