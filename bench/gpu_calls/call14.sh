R=bench/microbench/fir_row_test
# args: N U T S frames iters tg_log2 store ctas dbg pad spec
{
timeout 60 $R 51 1 16 64 1000 1 5 0 0 0 1 1; timeout 60 $R 51 2 8 67 1024 1 6 0 0 0 0 1; timeout 60 $R 101 2 16 67 520 1 5 2 0 0 0 1; timeout 60 $R 33 4 4 5 256 1 7 2 0 0 0 1; timeout 60 $R 101 4 8 33 1000 1 5 2 0 0 0 1
for spec in 0 1; do
timeout 60 $R 51 1 16 65536 1024 10 5 0 0 0 1 $spec
timeout 60 $R 101 1 16 32768 1024 10 5 0 0 0 1 $spec
timeout 60 $R 51 2 8 32768 1024 10 6 0 0 0 0 $spec
timeout 60 $R 101 2 16 32768 1024 10 5 2 0 0 0 $spec
timeout 60 $R 51 4 4 32768 1024 10 7 2 0 0 0 $spec
timeout 60 $R 101 4 8 32768 1024 10 5 2 0 0 0 $spec
timeout 60 $R 251 4 8 16384 1024 10 5 2 0 0 0 $spec
done
} > gpurun_out/row_probe7.txt 2>&1
grep -E "FAIL|rror|mismatch" gpurun_out/row_probe7.txt | head; grep -c "ok$" gpurun_out/row_probe7.txt; grep RESULT gpurun_out/row_probe7.txt | tail -14 | sed -E 's/  RESULT //; s/ ctas=[0-9] dbg=0//; s/TFLOP.*//'
for spec in 0 1; do ncu --metrics smsp__inst_executed.sum --clock-control none -k regex:fir_row -s 1 -c 1 $R 51 1 16 65536 1024 2 5 0 0 0 1 $spec 2>&1 | grep -E "inst_executed"; ncu --metrics smsp__inst_executed.sum --clock-control none -k regex:fir_row -s 1 -c 1 $R 51 2 8 32768 1024 2 6 0 0 0 0 $spec 2>&1 | grep -E "inst_executed"; done
