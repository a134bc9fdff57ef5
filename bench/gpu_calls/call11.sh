R=bench/microbench/fir_row_test
# args: N U T S frames iters tg_log2 store ctas dbg pad tiles
{
for st in 0 2; do for pad in 0 1; do
timeout 60 $R 51 4 4 64 1024 1 5 $st 0 0 $pad 2; timeout 60 $R 51 1 16 64 1000 1 5 $st 0 0 $pad 2; timeout 60 $R 101 2 8 67 1024 1 6 $st 0 0 $pad 2; timeout 60 $R 33 4 4 5 256 1 7 $st 0 0 $pad 2; timeout 60 $R 63 1 16 33 520 1 5 $st 0 0 $pad 2; timeout 60 $R 27 1 8 9 2048 1 5 $st 0 0 $pad 2; timeout 60 $R 251 2 8 9 4096 1 5 $st 0 0 $pad 2
done; done
for tiles in 1 2; do
timeout 60 $R 51 1 16 65536 1024 10 5 0 0 0 1 $tiles
timeout 60 $R 51 1 16 65536 1024 10 5 0 0 0 0 $tiles
timeout 60 $R 51 1 8 65536 1024 10 5 0 0 0 0 $tiles
timeout 60 $R 51 1 8 65536 1024 10 6 0 0 0 0 $tiles
timeout 60 $R 51 1 8 65536 1024 10 5 0 0 0 1 $tiles
timeout 60 $R 51 2 8 32768 1024 10 5 0 0 0 0 $tiles
timeout 60 $R 51 2 8 32768 1024 10 6 0 0 0 0 $tiles
timeout 60 $R 51 2 8 32768 1024 10 6 2 0 0 0 $tiles
timeout 60 $R 51 2 8 32768 1024 10 5 2 0 0 0 $tiles
timeout 60 $R 51 4 4 32768 1024 10 7 2 0 0 0 $tiles
timeout 60 $R 51 4 4 32768 1024 10 6 2 0 0 0 $tiles
timeout 60 $R 51 4 4 32768 1024 10 5 2 0 0 0 $tiles
timeout 60 $R 101 4 4 32768 1024 10 7 2 0 0 0 $tiles
timeout 60 $R 101 4 4 32768 1024 10 6 2 0 0 0 $tiles
timeout 60 $R 101 2 8 32768 1024 10 6 2 0 0 0 $tiles
timeout 60 $R 101 2 8 32768 1024 10 5 2 0 0 0 $tiles
timeout 60 $R 101 2 8 32768 1024 10 6 0 0 0 0 $tiles
timeout 60 $R 101 1 16 32768 1024 10 5 0 0 0 1 $tiles
timeout 60 $R 101 1 8 32768 1024 10 5 0 0 0 0 $tiles
timeout 60 $R 251 1 16 16384 1024 10 5 0 0 0 1 $tiles
timeout 60 $R 251 2 8 16384 1024 10 5 2 0 0 0 $tiles
timeout 60 $R 251 4 4 16384 1024 10 6 2 0 0 0 $tiles
done
} > gpurun_out/row_probe6.txt 2>&1
grep -E "FAIL|rror|mismatch" gpurun_out/row_probe6.txt | head; grep -c "ok$" gpurun_out/row_probe6.txt; grep RESULT gpurun_out/row_probe6.txt | tail -44 | sed -E 's/  RESULT //; s/ ctas=[0-9] dbg=0//; s/TFLOP.*//'
