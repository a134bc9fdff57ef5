python -m pytest tests/test_gpu_agc.py -q -x 2>&1 | tail -2
for cfg in "16384 1024 512" "65536 1024 512" "16384 512 256" "32768 256 128" "16384 2048 512"; do set -- $cfg; python bench/agc_perf.py --streams $1 --frames $2 --window $3 2>&1 | tail -1; done
