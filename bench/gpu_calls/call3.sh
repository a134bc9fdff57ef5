cd bench/microbench && ./store_pattern2 > ../../gpurun_out/store_pattern2.txt 2>&1; cd ../..
R=bench/microbench/fir_row_test_dbg
{
for cfg in "51 4 4 32768 6" "51 4 8 32768 6" "51 1 16 65536 5" "101 4 8 32768 5" "51 2 8 32768 5" "251 1 16 16384 5"; do set -- $cfg
for d in 0 1 2 4 3 6 7; do $R $1 $2 $3 $4 1024 10 $5 0 0 $d | grep RESULT; done; done
} > gpurun_out/row_phases.txt 2>&1
R=bench/microbench/fir_row_test
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 1 -c 1 -o gpurun_out/row_u4t4 $R 51 4 4 32768 1024 2 6 0 > gpurun_out/ncu_row1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 1 -c 1 -o gpurun_out/row_u1t16 $R 51 1 16 65536 1024 2 5 0 > gpurun_out/ncu_row2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fir_row -s 1 -c 1 -o gpurun_out/row_u4t8_n101 $R 101 4 8 32768 1024 2 5 0 > gpurun_out/ncu_row3.log 2>&1
cat gpurun_out/store_pattern2.txt gpurun_out/row_phases.txt
