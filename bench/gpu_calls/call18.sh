for mb in 64 16 8 4 2; do AFP_HOST_CHUNK_MB=$mb python bench/e2e_probe.py 8 2>&1 | tail -1; done
python -m pytest tests -m gpu -q -k "host or callback or compat or modes" 2>&1 | tail -3
