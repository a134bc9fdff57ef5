#!/bin/bash
# does the split-FP16 tensor kernel beat the FFMA2 single-row kernel below 400 taps?
for taps in 51 101 251 385; do for streams in 16384 65536; do for min in 5000 32; do
  echo -n "min_taps $min: "; AFP_TC_MIN_TAPS=$min python bench/quick_perf.py --taps $taps --up 1 --frames 1024 --streams $streams 2>&1 | tail -1 | cut -c1-220
done; done; done
AFP_TC_MIN_TAPS=32 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor2.py -x -q 2>&1 | tail -3
