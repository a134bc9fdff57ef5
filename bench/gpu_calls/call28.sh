python bench/agc_perf.py --streams 65536 --frames 1024 --window 512 --iters 4 > gpurun_out/agc_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:agc_kernel -s 3 -c 1 -o gpurun_out/agc_prof -f python bench/agc_perf.py --streams 65536 --frames 1024 --window 512 --iters 4 > gpurun_out/agc_ncu.log 2>&1
tail -2 gpurun_out/agc_ncu.log
python bench/callback_latency.py 2>&1 | tail -1
