#!/bin/bash
timeout 20 bench/microbench/tc2_probe 251 1024 512 4 2 | tail -2 || { echo "probe hung / failed"; exit 1; }
timeout 600 python -m pytest tests/test_gpu_tensor2.py tests/test_gpu_groups.py tests/test_gpu_configs_full_size.py tests/test_gpu_parity.py -x -q --timeout 90 2>&1 | tail -4
{
for t in 51 251 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 | tail -1; done
for t in 51 251 1001; do I24=1 timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 | tail -1; done
for t in 51 251 1001; do P24=1 timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 | tail -1; done
} > gpurun_out/tc2_probe_epi.txt 2>&1
cut -c1-130 gpurun_out/tc2_probe_epi.txt; cut -c300-460 gpurun_out/tc2_probe_epi.txt
