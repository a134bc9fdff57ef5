timeout 300 python -m pytest tests/test_gpu_tensor_path.py tests/test_gpu_live_swap.py -q -x 2>&1 | tail -4
timeout 120 python bench/quick_perf.py --taps 1001 --up 1 --frames 512 --streams 16384 --fmt i24 --dither tpdf --iters 20 | tail -1 | cut -c1-200
timeout 120 python bench/quick_perf.py --taps 1001 --up 1 --frames 512 --streams 16384 --iters 20 | tail -1 | cut -c1-200
timeout 120 python bench/quick_perf.py --taps 501 --up 1 --frames 1024 --streams 16384 --iters 20 | tail -1 | cut -c1-200
timeout 120 python bench/quick_perf.py --taps 701 --up 1 --frames 512 --streams 16384 --iters 20 | tail -1 | cut -c1-200
timeout 120 python bench/quick_perf.py --taps 1001 --up 1 --frames 1024 --streams 16384 --fmt i24 --dither tpdf --iters 20 | tail -1 | cut -c1-200
