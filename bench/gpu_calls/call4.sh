R=bench/microbench/fir_row_test
{
for st in 0 2; do
$R 51 4 8 64 1024 1 5 $st; $R 51 4 4 64 1024 1 5 $st; $R 51 1 16 64 1024 1 5 $st; $R 51 1 8 64 1000 1 5 $st; $R 101 2 8 67 1024 1 6 $st; $R 251 2 16 64 512 1 5 $st; $R 33 4 8 5 256 1 7 $st
$R 52 1 16 33 1000 1 5 $st; $R 63 1 16 33 520 1 5 $st; $R 64 1 16 33 1024 1 6 $st; $R 7 1 16 9 1024 1 5 $st; $R 27 2 8 9 1000 1 5 $st; $R 99 4 8 9 1024 1 5 $st; $R 103 4 4 9 1024 1 5 $st; $R 45 2 16 9 2048 1 5 $st
done
for st in 0 2; do for tg in 5 6; do
$R 51 4 8 32768 1024 10 $tg $st; $R 51 4 4 32768 1024 10 $tg $st
$R 101 4 8 32768 1024 10 $tg $st
$R 51 1 16 65536 1024 10 $tg $st; $R 51 1 8 65536 1024 10 $tg $st
$R 51 2 8 32768 1024 10 $tg $st; $R 51 2 16 32768 1024 10 $tg $st
$R 101 1 16 32768 1024 10 $tg $st; $R 101 2 8 32768 1024 10 $tg $st
$R 251 1 16 16384 1024 10 $tg $st; $R 251 4 8 16384 1024 10 $tg $st
done; done
} > gpurun_out/row_probe2.txt 2>&1
R=bench/microbench/fir_row_test_dbg
{
for cfg in "51 4 8 32768 5" "51 1 16 65536 5"; do set -- $cfg
for d in 0 1 2 4 3 6 7; do $R $1 $2 $3 $4 1024 10 $5 2 0 $d | grep RESULT; done; done
} > gpurun_out/row_phases2.txt 2>&1
R=bench/microbench/fir_row_test
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 1 -c 1 -o gpurun_out/row2_u4t8 $R 51 4 8 32768 1024 2 5 2 > gpurun_out/ncu_row1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 1 -c 1 -o gpurun_out/row2_u1t16 $R 51 1 16 65536 1024 2 5 2 > gpurun_out/ncu_row2.log 2>&1
grep -E "FAIL|rror|mismatch" gpurun_out/row_probe2.txt | head; grep -c "ok$" gpurun_out/row_probe2.txt; grep RESULT gpurun_out/row_probe2.txt | tail -44; cat gpurun_out/row_phases2.txt
