#!/bin/bash
# round 2, call 2: first run of the split-FP16 tensor kernel: tests, then timing
timeout 600 python -m pytest tests/test_gpu_tensor2.py -x -q > gpurun_out/r2c2_pytest_tc2.txt 2>&1; echo "tc2 tests rc=$?"
tail -15 gpurun_out/r2c2_pytest_tc2.txt
timeout 600 python -m pytest tests/test_gpu_tensor_path.py -x -q > gpurun_out/r2c2_pytest_tc.txt 2>&1; echo "tensor tests rc=$?"
tail -5 gpurun_out/r2c2_pytest_tc.txt
timeout 300 python bench/tc2_perf.py > gpurun_out/r2c2_perf.txt 2>&1; echo "perf rc=$?"
cat gpurun_out/r2c2_perf.txt
