timeout 300 python -m pytest tests/test_gpu_spectrum.py -q -x 2>&1 | tail -12
for cfg in "16384 1024" "65536 1024" "16384 4096" "65536 256"; do set -- $cfg; timeout 100 python bench/spectrum_perf.py --streams $1 --frames $2 | tail -1; done
