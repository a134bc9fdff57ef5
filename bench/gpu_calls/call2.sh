# row-kernel probe: correctness on small S, then timing on the large cases (direct vs bulk stores, tile sizes, group sizes)
cd bench/microbench && ./ffma_peak 2>&1 | grep -E "x2|ur" > ../../gpurun_out/ffma_peak3.jsonl; cd ../..
R=bench/microbench/fir_row_test
{
for st in 0 1; do
$R 51 4 8 64 1024 1 5 $st; $R 51 4 4 64 1024 1 5 $st; $R 51 1 16 64 1024 1 5 $st; $R 51 1 8 64 1000 1 4 $st; $R 101 2 8 67 1024 1 6 $st; $R 251 2 16 64 512 1 5 $st; $R 33 4 8 5 256 1 7 $st
done
for st in 0 1; do for tg in 5 6 7; do
$R 51 4 8 32768 1024 10 $tg $st; $R 51 4 4 32768 1024 10 $tg $st
$R 101 4 8 32768 1024 10 $tg $st
$R 51 1 16 65536 1024 10 $tg $st; $R 51 1 8 65536 1024 10 $tg $st
$R 51 2 8 32768 1024 10 $tg $st; $R 51 2 16 32768 1024 10 $tg $st
$R 101 1 16 32768 1024 10 $tg $st; $R 101 2 8 32768 1024 10 $tg $st
$R 251 1 16 16384 1024 10 $tg $st; $R 251 4 8 16384 1024 10 $tg $st
done; done
} > gpurun_out/row_probe1.txt 2>&1
grep -E "RESULT|FAIL|error" gpurun_out/row_probe1.txt | tail -80
