python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_r1e.txt; cat gpurun_out/pytest_gpu_r1e.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 3000 gpurun_out/bench10.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 2 --cpu-seconds 1 > gpurun_out/ncu_b10.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 3 -c 1 -o gpurun_out/r01b_row_n51u4 python bench/quick_perf.py --taps 51 --up 4 --frames 1024 --streams 32768 --iters 2 > gpurun_out/ncu_q1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 3 -c 1 -o gpurun_out/r01b_row_n51u1 python bench/quick_perf.py --taps 51 --up 1 --frames 1024 --streams 65536 --iters 2 > gpurun_out/ncu_q2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 3 -c 1 -o gpurun_out/r01b_row_n101u2 python bench/quick_perf.py --taps 101 --up 2 --frames 1024 --streams 32768 --iters 2 > gpurun_out/ncu_q3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
