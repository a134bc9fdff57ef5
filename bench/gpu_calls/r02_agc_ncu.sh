#!/bin/bash
CMD="python bench/agc_perf.py --streams ${AGC_S:-16384} --frames 1024 --window 512 --iters 4"
$CMD > gpurun_out/agc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:agc_kernel -s 3 -c 1 -o gpurun_out/r02_agc $CMD > gpurun_out/agc_ncu.log 2>&1
tail -2 gpurun_out/agc_ncu.log
