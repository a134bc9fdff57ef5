#!/bin/bash
# round 2, call 6: history copy in the epilogue warps, fresh windows in one round trip of the split-FP16 tensor kernel
P=bench/microbench/tc2_probe
for args in "1001 16384 512 16 10" "1001 16384 512 1 10" "1001 16384 512 2 10" "1001 16384 512 3 10" "513 16384 512 16 10" "449 16384 512 16 10" "1001 16384 4096 2 10" "1001 65536 512 16 5"; do
  timeout 120 $P $args | tail -2
done > gpurun_out/r2c6_probe.txt 2>&1
I24=1 timeout 120 $P 1001 16384 512 16 10 | tail -1 >> gpurun_out/r2c6_probe.txt 2>&1
cat gpurun_out/r2c6_probe.txt
timeout 600 python -m pytest tests/test_gpu_tensor2.py tests/test_gpu_tensor_path.py -x -q 2>&1 | tail -15
