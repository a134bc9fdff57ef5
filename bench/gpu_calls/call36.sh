python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench/agc_perf.py --streams 16384 --frames 1024 --window 512 --iters 4 > gpurun_out/agc_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:agc_kernel -s 3 -c 1 -f -o gpurun_out/r01_agc_16384x1024 python bench/agc_perf.py --streams 16384 --frames 1024 --window 512 --iters 4 > gpurun_out/ncu_f4.log 2>&1
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'],d['clocks'])
for r in d['agc_kernel_only']: print('agc', r)
PY
