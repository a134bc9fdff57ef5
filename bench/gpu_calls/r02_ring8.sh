#!/bin/bash
for cfg in "251 64 512 2 1 2" "1001 40 512 3 1 2" "251 64 512 2 1 4" "1001 40 1024 2 1 4" "96 33 512 1 1 2" "513 40 512 3 1 1" "514 40 512 3 1 1" "251 1024 512 4 2 1"; do
  timeout 25 bench/microbench/tc2_probe $cfg | grep -E "max abs|mismatch|error" | head -4 || echo "probe failed: $cfg"
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -5
{
for t in 51 251 449 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 1 | tail -1; done
for u in 2 4; do for t in 251 501 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 $u | tail -1; done; done
} > gpurun_out/tc2_ring8_probe.txt 2>&1
cut -c1-330 gpurun_out/tc2_ring8_probe.txt
