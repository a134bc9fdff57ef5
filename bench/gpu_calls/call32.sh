python -m pytest tests/test_gpu_agc.py -q -x 2>&1 | tail -3
AFP_AGC_TEAM=1 python -m pytest tests/test_gpu_agc.py -q -x 2>&1 | tail -1
for cfg in "16384 1024 512" "65536 1024 512" "4096 1024 512" "16384 512 256" "32768 256 128" "16384 2048 512" "8192 4096 512"; do set -- $cfg; for t in 1 2; do echo -n "team $t: "; AFP_AGC_TEAM=$t python bench/agc_perf.py --streams $1 --frames $2 --window $3 2>&1 | tail -1; done; done
