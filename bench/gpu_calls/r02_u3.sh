#!/bin/bash
for cfg in "251 64 512 2 1 3" "1001 40 512 3 1 3" "90 33 1024 1 1 3" "251 64 512 2 1 2" "251 64 512 2 1 4"; do
  timeout 25 bench/microbench/tc2_probe $cfg | grep -E "max abs|mismatch|error" | head -4 || echo "probe failed: $cfg"
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4
{
for t in 101 251 501 1001; do timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 3 | tail -1; done
for t in 51 101 129 251 1001; do echo "FFMA2 taps=$t U=3: $(AFP_NO_TC2_UP=1 python bench/quick_perf.py --taps $t --up 3 --frames 1024 --streams 65536 --iters 10 2>&1 | tail -1 | cut -c1-150)"; echo "tensor taps=$t U=3: $(AFP_TC_MIN_TAPS=48 python bench/quick_perf.py --taps $t --up 3 --frames 1024 --streams 65536 --iters 10 2>&1 | tail -1 | cut -c1-150)"; done
} > gpurun_out/tc2_u3_probe.txt 2>&1
cut -c1-260 gpurun_out/tc2_u3_probe.txt
