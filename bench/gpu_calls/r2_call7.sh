#!/bin/bash
# round 2, call 7: the new bench.py end to end (driver's command line) and the whole GPU test suite
( time timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c7_bench.err
head -c 3000 gpurun_out/r2c7_bench.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c7_pytest.txt 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c7_pytest.txt
