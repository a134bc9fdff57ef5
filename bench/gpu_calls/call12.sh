python -m pytest tests -m gpu -q 2>&1 | tail -8
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 10 $5 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['stage'][-1]; print('U=$1 N=$2 B=$3 S=$4 $5', d['gsamples_s'], 'Gs/s', d['alg_gbs'], 'GB/s', 'kernel', st.get('kernel'))"; }
{
run 4 51 1024 32768 "--fmt i24 --dither tpdf"; run 4 51 1024 32768 "--fmt i16 --dither tpdf"; run 4 51 1024 32768 "--fmt i24_packed --dither tpdf"
run 1 51 1024 65536 "--fmt i16 --dither tpdf"; run 1 51 1024 65536 "--fmt i24 --dither tpdf"; run 2 51 1024 32768 "--fmt i24 --dither tpdf"
run 1 1 1024 65536 "--fmt i24 --dither tpdf"; run 1 1 1024 65536 "--fmt i16 --dither tpdf"; run 1 1 1024 65536 "--fmt i24"; run 1 1 1024 65536 ""; run 1 1 4096 16384 "--fmt i24 --dither tpdf"
run 1 1001 512 16384 "--fmt i24 --dither tpdf"; run 1 1001 512 16384 ""
} > gpurun_out/sweep_dither.txt 2>&1
cat gpurun_out/sweep_dither.txt
