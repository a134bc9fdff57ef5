#!/bin/bash
for i in 1 2; do timeout 30 bench/microbench/tc2_probe 1001 16384 512 16 5 2 | tail -1 | cut -c1-420; done
NOSTAGE=1 timeout 30 bench/microbench/tc2_probe 1001 16384 512 16 5 2 | tail -1 | cut -c1-420
timeout 30 bench/microbench/tc2_probe 897 16384 512 16 5 2 | tail -1 | cut -c1-420
timeout 30 bench/microbench/tc2_probe 1001 16384 512 16 20 2 | tail -1 | cut -c1-420
timeout 30 bench/microbench/tc2_probe 513 16384 512 16 5 1 | tail -1 | cut -c1-420
python bench/quick_perf.py --taps 1001 --up 2 --frames 1024 --streams 16384 --iters 10 2>&1 | tail -1 | cut -c1-300
