#!/bin/bash
# time-chunked AGC kernel: parity, then throughput over batch / block sizes and group sizes
python -m pytest tests/test_gpu_agc.py tests/test_gpu_agc_exact_gain.py -q -x 2>&1 | tail -3
for cfg in "16384 1024 512" "65536 1024 512" "4096 1024 512" "16384 512 256" "32768 256 128" "16384 2048 512" "8192 4096 512"; do
  set -- $cfg
  for g in ${GROUPS_TO_TRY:-0 8 16 22 32}; do
    echo -n "group $g: "; AFP_AGC_GROUP=$g python bench/agc_perf.py --streams $1 --frames $2 --window $3 2>&1 | tail -1
  done
done
