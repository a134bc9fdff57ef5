#!/bin/bash
for frames in 512 1024; do for taps in 128 160 192 224; do for min in 5000 32; do
  echo -n "frames $frames min_taps $min: "; AFP_TC_MIN_TAPS=$min python bench/quick_perf.py --taps $taps --up 1 --frames $frames --streams 16384 2>&1 | tail -1 | cut -c1-130
done; done; done
