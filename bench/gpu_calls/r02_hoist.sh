#!/bin/bash
timeout 20 bench/microbench/tc2_probe 251 1024 512 4 2 | grep "max abs" || { echo "probe hung / failed"; exit 1; }
timeout 900 python -m pytest tests/test_gpu_tensor2.py tests/test_gpu_groups.py tests/test_gpu_configs_full_size.py tests/test_gpu_smoke.py -x -q --timeout 120 2>&1 | tail -3
{
for t in 51 101 251 449 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 1 | tail -1; done
for t in 51 251 1001; do timeout 30 bench/microbench/tc2_probe $t 65536 1024 1 5 1 | tail -1; done
for u in 2 4; do for t in 251 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 $u | tail -1; done; done
P24=1 timeout 30 bench/microbench/tc2_probe 1001 16384 512 16 5 1 | tail -1
} > gpurun_out/tc2_hoist_probe.txt 2>&1
cut -c1-300 gpurun_out/tc2_hoist_probe.txt
