python -m pytest tests -m gpu -q 2>&1 | tail -6
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 10 $5 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('U=$1 N=$2 B=$3 S=$4 $5 NOCHAIN=$AFP_NO_TC_CHAIN', d['gsamples_s'], 'Gs/s', d['ms'], 'ms', d['tflops'], 'TF', [ (s['kernel'], s['n_taps']) for s in d['stage']])"; }
run 1 1001 4096 4096 ""; run 1 1001 1024 16384 "--fmt i24 --dither tpdf"; run 1 641 2048 8192 ""; run 1 449 4096 4096 ""
export AFP_NO_TC_CHAIN=1
run 1 1001 4096 4096 ""; run 1 1001 1024 16384 "--fmt i24 --dither tpdf"; run 1 641 2048 8192 ""; run 1 449 4096 4096 ""
