#!/bin/bash
# split-FP16 kernel with the N=128 [hi|lo] product: sanity, parity tests, probe counters (full lines into gpurun_out/)
timeout 20 bench/microbench/tc2_probe 251 1024 512 4 2 | tail -2 || { echo "probe hung / failed"; exit 1; }
timeout 400 python -m pytest tests/test_gpu_tensor2.py tests/test_gpu_groups.py tests/test_gpu_tensor_path.py tests/test_gpu_configs_full_size.py -x -q --timeout 90 2>&1 | tail -4
{
for t in 51 101 251 449 640 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 | tail -2; done
timeout 30 bench/microbench/tc2_probe 251 16384 512 2 5 | tail -1
timeout 30 bench/microbench/tc2_probe 1001 16384 512 1 5 | tail -1
timeout 30 bench/microbench/tc2_probe 1001 65536 512 16 5 | tail -1
} > gpurun_out/tc2_wide_probe.txt 2>&1
cut -c1-400 gpurun_out/tc2_wide_probe.txt
