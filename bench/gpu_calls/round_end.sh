# Round-end evidence run (one GPU): tests, smoke, bench, launch list, ncu captures of the three kernels.
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 2 --cpu-seconds 1 > gpurun_out/ncu_final.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_tc -s 2 -c 1 -o gpurun_out/r01_c2_bench_tc python bench.py --steps 1 --warmup 3 --blocks 2 --e2e-steps 2 --cpu-seconds 1 --no-sweep > gpurun_out/ncu_f0.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 3 -c 1 -o gpurun_out/r01_row_n51u4 python bench/quick_perf.py --taps 51 --up 4 --frames 1024 --streams 32768 --iters 2 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 3 -c 1 -o gpurun_out/r01_row_n51u1 python bench/quick_perf.py --taps 51 --up 1 --frames 1024 --streams 65536 --iters 2 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 3 -c 1 -o gpurun_out/r01_row_n101u2 python bench/quick_perf.py --taps 101 --up 2 --frames 1024 --streams 32768 --iters 2 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:agc_kernel -s 3 -c 1 -o gpurun_out/r01_agc_16384x1024 python bench/agc_perf.py --streams 16384 --frames 1024 --window 512 --iters 4 > gpurun_out/ncu_f4.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'],d['clocks'])
for r in d['c5_sweep_kernel_only']: print(r['taps'],r['upsample'],r['gsamples_s'],r['bound'],r['frac'])
for r in d['agc_kernel_only']: print('agc', r)
print(open('gpurun_out/bench_final_ref.json').read()[:300])
PY
ls -la gpurun_out/r01_*.ncu-rep
