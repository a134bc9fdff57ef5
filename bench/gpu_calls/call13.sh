python -m pytest tests -m gpu -q 2>&1 | tail -8
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 10 $5 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('U=$1 N=$2 B=$3 S=$4 $5 NO_COLLAPSE=$AFP_NO_COLLAPSE', d['gsamples_s'], 'Gs/s(out)', d['ms'], 'ms', [ (s['kernel'], s['n_taps'], s['upsample']) for s in d['stage']])"; }
{
run 4 51 1024 32768 "--fmt i24 --dither tpdf"; run 1 51 1024 65536 "--fmt i16 --dither tpdf"; run 1 1 1024 65536 "--fmt i24 --dither tpdf"; run 1 1001 512 16384 "--fmt i24 --dither tpdf"
run 4 251 1024 8192 "--chain reference"; run 2 251 1024 8192 "--chain reference"; run 4 1001 1024 4096 "--chain reference"
export AFP_NO_COLLAPSE=1
run 4 251 1024 8192 "--chain reference"; run 2 251 1024 8192 "--chain reference"; run 4 1001 1024 4096 "--chain reference"
} > gpurun_out/sweep_collapse.txt 2>&1
cat gpurun_out/sweep_collapse.txt
