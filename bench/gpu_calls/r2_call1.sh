#!/bin/bash
# round 2, call 1: tcgen05 issue-rate probe + state of the GPU test suite
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 300 bench/microbench/umma_peak 2.0 > gpurun_out/r2c1_umma.txt 2>&1; echo "umma rc=$?"
cat gpurun_out/r2c1_umma.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c1_pytest.txt
