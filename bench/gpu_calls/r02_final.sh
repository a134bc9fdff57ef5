#!/bin/bash
# Round-2 final evidence (one GPU): tests, smoke, both bench arms, launch list, ncu --set full of the headline kernel and of the polyphase form
python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; echo "ref rc=$?"
CMD="python bench.py --steps 5 --warmup 3 --no-sweep --no-configs --sustain-seconds 0 --cpu-seconds 1 --e2e-steps 2"
$CMD > gpurun_out/r02f_plain.log 2> gpurun_out/r02f_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 3 -c 1 -f -o gpurun_out/r02f_tc2_c2 $CMD > gpurun_out/r02f_ncu2.log 2>&1
tail -2 gpurun_out/r02f_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 2 -c 1 -f -o gpurun_out/r02f_tc2_poly bench/microbench/tc2_probe 251 16384 512 16 3 2 > gpurun_out/r02f_ncu3.log 2>&1
tail -2 gpurun_out/r02f_ncu3.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'sustained',d['value_sustained']['value'],'e2e',d['e2e']['value'],d['e2e'].get('frac'),'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'],d['clocks'])
for k,v in d['configs'].items(): print(k, v['value'], v['e2e']['value'], v['roofline'].get('bound'), v['roofline'].get('frac'))
for r in d['agc_kernel_only']+d['waterfall_rows_kernel_only']: print(r)
print(open('gpurun_out/bench_final_ref.json').read()[:400])
PY
