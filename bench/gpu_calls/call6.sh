R=bench/microbench/fir_row_test
{
for st in 0 2; do
timeout 60 $R 51 4 8 64 1024 1 5 $st 0 0 1; timeout 60 $R 51 1 16 64 1024 1 5 $st 0 0 1; timeout 60 $R 51 1 8 64 1000 1 5 $st 0 0 1; timeout 60 $R 101 2 8 67 1024 1 6 $st 0 0 1; timeout 60 $R 251 2 16 64 512 1 5 $st 0 0 1; timeout 60 $R 33 4 8 5 256 1 7 $st 0 0 1
timeout 60 $R 52 1 16 33 1000 1 5 $st 0 0 1; timeout 60 $R 63 1 16 33 520 1 5 $st 0 0 1; timeout 60 $R 64 1 16 33 1024 1 6 $st 0 0 1; timeout 60 $R 7 1 16 9 1024 1 5 $st 0 0 1; timeout 60 $R 27 2 8 9 1000 1 5 $st 0 0 1; timeout 60 $R 45 2 16 9 2048 1 5 $st 0 0 1
done
for pad in 1 0; do for st in 0 2; do for tg in 5 6; do
timeout 60 $R 51 1 16 65536 1024 10 $tg $st 0 0 $pad; timeout 60 $R 51 1 8 65536 1024 10 $tg $st 0 0 $pad
timeout 60 $R 101 1 16 32768 1024 10 $tg $st 0 0 $pad; timeout 60 $R 251 1 16 16384 1024 10 $tg $st 0 0 $pad
timeout 60 $R 51 2 8 32768 1024 10 $tg $st 0 0 $pad; timeout 60 $R 51 2 16 32768 1024 10 $tg $st 0 0 $pad
timeout 60 $R 101 2 8 32768 1024 10 $tg $st 0 0 $pad; timeout 60 $R 101 2 16 32768 1024 10 $tg $st 0 0 $pad
done; done; done
for tg in 5 6 7; do timeout 60 $R 51 4 8 32768 1024 10 $tg 2 0 0 1; timeout 60 $R 101 4 8 32768 1024 10 $tg 2 0 0 1; timeout 60 $R 251 4 8 16384 1024 10 $tg 0 0 0 1; done
for tg in 7; do timeout 60 $R 51 4 4 32768 1024 10 $tg 2 0 0 0; timeout 60 $R 51 4 8 32768 1024 10 $tg 2 0 0 0;  timeout 60 $R 101 4 8 32768 1024 10 $tg 2 0 0 0; timeout 60 $R 101 4 4 32768 1024 10 $tg 2 0 0 0; timeout 60 $R 101 4 4 32768 1024 10 6 2 0 0 0; done
} > gpurun_out/row_probe4.txt 2>&1
grep -E "FAIL|rror|mismatch" gpurun_out/row_probe4.txt | head; grep -c "ok$" gpurun_out/row_probe4.txt; grep RESULT gpurun_out/row_probe4.txt | tail -n +25
