#!/bin/bash
# Tile-size experiment: same configuration with each compiled chunk size T (AFP_FORCE_T).
run() { # U N S Tlist
  for T in $4; do
    AFP_FORCE_T=$T python bench/quick_perf.py --taps $2 --up $1 --frames 1024 --streams $3 --iters 10 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('U=$1 N=$2 T=$T', d['gsamples_s'], 'Gs/s', d['tflops'], 'TF', d['alg_gbs'], 'GB/s', 'tg', d['stage'][0]['threads_per_segment'], 'smem', d['stage'][0]['smem_bytes'])"
  done
}
run 1 51 32768 "8 16"; run 1 101 32768 "8 16"; run 1 251 16384 "8 16"; run 1 501 8192 "8 16"
run 2 51 32768 "4 8"; run 2 101 32768 "4 8"; run 2 251 16384 "4 8"; run 2 501 8192 "4 8"
run 4 51 32768 "4 8"; run 4 101 32768 "4 8"; run 4 251 16384 "4 8"; run 4 501 8192 "4 8"; run 4 1001 8192 "4 8"
run 3 129 16384 "4"
