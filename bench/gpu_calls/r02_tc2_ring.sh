#!/bin/bash
# every command under its own short timeout: a deadlocked kernel must not eat the call
timeout 20 bench/microbench/tc2_probe 251 1024 512 4 2 | tail -1 || { echo "probe hung / failed"; exit 1; }
timeout 240 python -m pytest tests/test_gpu_tensor2.py tests/test_gpu_groups.py tests/test_gpu_tensor_path.py -x -q --timeout 60 2>&1 | tail -4
for t in 51 251 449 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 | tail -1; done
timeout 30 bench/microbench/tc2_probe 251 16384 512 2 5 | tail -1
timeout 30 bench/microbench/tc2_probe 1001 16384 512 1 5 | tail -1
for taps in 51 101 251 385 1001; do echo -n "min 32: "; AFP_TC_MIN_TAPS=32 timeout 60 python bench/quick_perf.py --taps $taps --up 1 --frames 1024 --streams 16384 2>&1 | tail -1 | cut -c1-130; done
