#!/bin/bash
# polyphase interpolators on the tensor kernel: stand-alone check first (short timeout: a wrong barrier would hang), then tests, then speed
for cfg in "251 64 512 2 1 2" "1001 40 512 3 1 2" "251 64 512 2 1 4" "1001 40 1024 2 1 4" "96 33 512 1 1 2"; do
  timeout 25 bench/microbench/tc2_probe $cfg | grep -E "max abs|mismatch|error" | head -4 || echo "probe failed: $cfg"
done
timeout 25 bench/microbench/tc2_probe 251 1024 512 4 2 1 | grep -E "max abs" 
timeout 600 python -m pytest tests/test_gpu_tensor2.py -x -q --timeout 90 2>&1 | tail -6
{
for u in 2 4; do for t in 101 251 501 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 $u | tail -1; done; done
for u in 2 4; do for t in 251 1001; do timeout 30 bench/microbench/tc2_probe $t 65536 1024 1 5 $u | tail -1; done; done
} > gpurun_out/tc2_poly_probe.txt 2>&1
cut -c1-330 gpurun_out/tc2_poly_probe.txt
