for mb in 64 16 8 4 2 1; do AFP_HOST_CHUNK_MB=$mb python bench/e2e_probe.py 8 2>&1 | tail -1; done
python bench/microbench/pcie_bw.py 2>&1 | tail -1
python -m pytest tests -m gpu -q -k "host" 2>&1 | tail -3
