python -m pytest tests -m gpu -q -x 2>&1 | tail -8
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 10 $5 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('U=$1 N=$2 B=$3 S=$4 $5 NO_ROW=$AFP_NO_ROW', d['gsamples_s'], 'Gs/s', d['tflops'], 'TF', d['alg_gbs'], 'GB/s', [ (s['kernel'], s['n_taps'], s['tile_taps']) for s in d['stage']])"; }
run 3 51 1024 32768 ""; run 3 129 1024 16384 ""; run 3 251 1024 16384 ""; run 3 1001 1024 8192 ""; run 3 129 1024 16384 "--fmt i24 --dither tpdf"
export AFP_NO_ROW=1
run 3 51 1024 32768 ""; run 3 129 1024 16384 ""; run 3 251 1024 16384 ""; run 3 1001 1024 8192 ""; run 3 129 1024 16384 "--fmt i24 --dither tpdf"
