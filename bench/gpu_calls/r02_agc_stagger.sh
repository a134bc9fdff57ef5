#!/bin/bash
for cfg in "16384 1024 512" "65536 1024 512" "16384 512 256"; do
  set -- $cfg
  for st in 0 200 400 800 1600; do
    echo -n "stagger $st: "; AFP_AGC_STAGGER_NS=$st python bench/agc_perf.py --streams $1 --frames $2 --window $3 2>&1 | tail -1
  done
done
