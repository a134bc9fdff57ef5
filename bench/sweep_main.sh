#!/bin/bash
# Kernel-only throughput over the C5-style grid (f32 out), one line per case.
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 10 $5 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('U=$1 N=$2 B=$3 S=$4 $5', d['gsamples_s'], 'Gs/s', d['tflops'], 'TF', d['alg_gbs'], 'GB/s')"; }
run 1 1001 512 16384; run 1 1001 512 16384 "--fmt i24 --dither tpdf"
run 1 251 1024 16384; run 1 51 1024 65536
run 2 51 1024 32768; run 2 251 1024 16384; run 2 1001 1024 8192
run 4 51 1024 32768; run 4 101 1024 32768; run 4 251 1024 16384; run 4 1001 1024 8192
run 3 129 1024 16384
run 4 51 1024 32768 "--fmt i24 --dither tpdf"
