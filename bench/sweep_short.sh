#!/bin/bash
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 10 $5 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('U=$1 N=$2 B=$3 S=$4 $5', d['gsamples_s'], 'Gs/s', d['tflops'], 'TF', d['alg_gbs'], 'GB/s')"; }
run 1 51 1024 65536; run 2 51 1024 32768; run 4 51 1024 32768; run 4 101 1024 32768; run 1 101 1024 32768; run 2 101 1024 32768
