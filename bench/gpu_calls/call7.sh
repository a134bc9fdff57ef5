python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1c.txt; cat gpurun_out/pytest_gpu_r1c.txt
run() { python bench/quick_perf.py --taps $2 --up $1 --frames $3 --streams $4 --iters 10 $5 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['stage'][-1] if isinstance(d['stage'],list) else d['stage']; print('U=$1 N=$2 B=$3 S=$4 $5 NO_ROW=$AFP_NO_ROW', d['gsamples_s'], 'Gs/s', d['tflops'], 'TF', d['alg_gbs'], 'GB/s', 'kernel', st.get('kernel'), 'T', st.get('tile_taps'), 'tg', st.get('threads_per_segment'))"; }
{
for nr in "" 1; do export AFP_NO_ROW=$nr
run 1 51 1024 65536; run 2 51 1024 32768; run 4 51 1024 32768; run 4 101 1024 32768; run 1 101 1024 32768; run 2 101 1024 32768
run 1 251 1024 16384; run 2 251 1024 16384; run 4 251 1024 16384; run 2 501 1024 8192; run 4 501 1024 8192; run 4 1001 1024 8192; run 1 501 1024 8192
run 4 51 1024 32768 "--fmt i24 --dither tpdf"; run 1 51 1024 65536 "--fmt i16 --dither tpdf"; run 4 51 512 32768; run 4 51 256 65536
done
} > gpurun_out/sweep_row_lib.txt 2>&1
cat gpurun_out/sweep_row_lib.txt
