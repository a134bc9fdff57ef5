python bench/callback_latency.py --cpu 2>&1 | tail -1
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 20 $5 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('U=$1 N=$2 B=$3 S=$4 $5', d['gsamples_s'], 'Gs/s', d['ms'], 'ms', d['tflops'], 'TF', [ (s['kernel'], s['n_taps'], s['upsample']) for s in d['stage']])"; }
run 2 251 1024 4096 ""; run 2 251 1024 4096 "--fmt i24 --dither tpdf"; run 4 251 1024 1024 "--chain reference"; run 4 251 1024 65536 "--chain reference"; run 3 129 1024 16384 ""
run 1 1001 512 1024 "--fmt i24 --dither tpdf"; run 1 1001 512 4096 "--fmt i24 --dither tpdf"; run 1 1001 512 65536 "--fmt i24 --dither tpdf"
