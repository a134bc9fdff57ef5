#!/bin/bash
P24=1 timeout 120 ncu --set full --import-source on --clock-control none -k regex:fir_tc2 -s 1 -c 1 -o gpurun_out/r02_tc2_p24_51 -f bench/microbench/tc2_probe 51 16384 512 16 2 > gpurun_out/ncu_epi.log 2>&1
tail -3 gpurun_out/ncu_epi.log
