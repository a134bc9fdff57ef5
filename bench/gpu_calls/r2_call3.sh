#!/bin/bash
# round 2, call 3: where the split-FP16 tensor kernel's time goes (probe counters, ncu), tests after the refactor
P=bench/microbench/tc2_probe
for args in "1001 16384 512 16 10" "1001 16384 512 1 10" "1001 2368 512 16 10" "513 16384 512 16 10" "449 16384 512 16 10" "1001 16384 4096 2 10"; do
  timeout 120 $P $args | tail -2
done > gpurun_out/r2c3_probe.txt 2>&1
I24=1 timeout 120 $P 1001 16384 512 16 10 | tail -1 >> gpurun_out/r2c3_probe.txt 2>&1
cat gpurun_out/r2c3_probe.txt
timeout 600 python -m pytest tests/test_gpu_tensor2.py -x -q 2>&1 | tail -3
timeout 120 $P 1001 16384 512 16 3 > gpurun_out/r2c3_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 1 -c 1 -o gpurun_out/r2c3_tc2 $P 1001 16384 512 16 3 > gpurun_out/r2c3_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2c3_ncu.log
