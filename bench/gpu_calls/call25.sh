python -m pytest tests -m gpu -q 2>&1 | tail -6
for a in "501 301 8192 1024" "129 201 16384 1024" "501 501 8192 512" "51 101 32768 1024"; do
python bench/c3_perf.py $a 2>&1 | tail -1; AFP_NO_COLLAPSE=1 python bench/c3_perf.py $a 2>&1 | tail -1
done
