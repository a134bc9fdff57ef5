set -x
cd bench/microbench && ./ffma_peak > ../../gpurun_out/ffma_peak2.jsonl 2>&1; cd ../..
for cfg in "51 4 32768" "51 1 65536" "101 4 32768" "51 2 32768"; do set -- $cfg
for d in 0 1 2 4 3 5 6 7; do
echo "N=$1 U=$2 S=$3 DBG=$d $(AFP_DBG=$d python bench/quick_perf.py --taps $1 --up $2 --frames 1024 --streams $3 --iters 10 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms'],'ms',d['gsamples_s'],'Gs/s')")"
done; done > gpurun_out/dbg_phases.txt 2>&1
