python -m pytest tests/test_gpu_wire_format.py tests/test_gpu_agc.py tests/test_gpu_modes.py -q -x 2>&1 | tail -8
python bench/agc_perf.py --streams 16384 --frames 1024 --window 512 | tail -1
AFP_AGC_3CTA=1 python bench/agc_perf.py --streams 16384 --frames 1024 --window 512 | tail -1
AFP_AGC_3CTA=1 python bench/agc_perf.py --streams 65536 --frames 1024 --window 512 | tail -1
AFP_AGC_3CTA=1 python bench/agc_perf.py --streams 16384 --frames 512 --window 256 | tail -1
python bench/callback_latency.py 2>&1 | tail -1
