R=bench/microbench/fir_row_test
{
for st in 0 2; do for pad in 0 1; do
timeout 60 $R 51 4 8 64 1024 1 5 $st 0 0 $pad; timeout 60 $R 51 1 16 64 1000 1 5 $st 0 0 $pad; timeout 60 $R 101 2 8 67 1024 1 6 $st 0 0 $pad; timeout 60 $R 33 4 4 5 256 1 7 $st 0 0 $pad; timeout 60 $R 63 1 16 33 520 1 5 $st 0 0 $pad
done; done
timeout 60 $R 51 1 16 65536 1024 10 5 0 0 0 1
timeout 60 $R 51 2 8 32768 1024 10 6 0 0 0 0
timeout 60 $R 51 2 8 32768 1024 10 7 0 0 0 0
timeout 60 $R 51 4 4 32768 1024 10 7 2 0 0 0
timeout 60 $R 101 4 8 32768 1024 10 5 2 0 0 0
timeout 60 $R 101 4 8 32768 1024 10 7 2 0 0 0
timeout 60 $R 101 2 16 32768 1024 10 5 2 0 0 0
timeout 60 $R 101 2 16 32768 1024 10 6 2 0 0 0
timeout 60 $R 101 1 16 32768 1024 10 5 0 0 0 1
timeout 60 $R 251 1 16 16384 1024 10 5 0 0 0 1
} > gpurun_out/row_probe5.txt 2>&1
grep -E "FAIL|rror|mismatch" gpurun_out/row_probe5.txt | head; grep -c "ok$" gpurun_out/row_probe5.txt; grep RESULT gpurun_out/row_probe5.txt | tail -10
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:fir_row -s 1 -c 1 $R 51 1 16 65536 1024 2 5 0 0 0 1 2>&1 | grep -E "inst_executed|duration"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:fir_row -s 1 -c 1 $R 51 2 8 32768 1024 2 6 0 0 0 0 2>&1 | grep -E "inst_executed|duration"
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
