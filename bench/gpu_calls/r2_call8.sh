#!/bin/bash
# round 2, call 8: full-size config tests (C3 10x501, C1 431 blocks, full-scale FFMA2, wire format vs oracle)
timeout 900 python -m pytest tests/test_gpu_configs_full_size.py -x -q > gpurun_out/r2c8_pytest.txt 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r2c8_pytest.txt
