run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 20 $5 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('U=$1 N=$2 B=$3 S=$4 $5 TCMIN=$AFP_TC_MIN_TAPS', d['gsamples_s'], 'Gs/s', d['ms'], 'ms', d['tflops'], 'TF', [ (s['kernel'], s['n_taps'], s['upsample']) for s in d['stage']])"; }
for n in 385 449 501 577 641; do
export AFP_TC_MIN_TAPS=300; run 1 $n 512 16384 ""; run 1 $n 1024 16384 ""
export AFP_TC_MIN_TAPS=5000; run 1 $n 512 16384 ""; run 1 $n 1024 16384 ""
done
