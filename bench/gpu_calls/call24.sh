R=bench/microbench/fir_row_test
# args: N U T S frames iters tg_log2 store ctas dbg pad spec NT
{
timeout 60 $R 51 1 16 64 1000 1 5 0 0 0 1 1 128; timeout 60 $R 51 2 8 67 1024 1 6 0 0 0 0 1 128
for nt in 256 128; do
timeout 60 $R 51 1 16 65536 1024 10 5 0 0 0 1 1 $nt
timeout 60 $R 101 1 16 32768 1024 10 5 0 0 0 1 1 $nt
timeout 60 $R 251 1 16 16384 1024 10 5 0 0 0 1 1 $nt
timeout 60 $R 51 2 8 32768 1024 10 6 0 0 0 0 1 $nt
timeout 60 $R 101 2 8 32768 1024 10 6 0 0 0 0 1 $nt
done
} > gpurun_out/row_probe8.txt 2>&1
grep -E "FAIL|rror|mismatch" gpurun_out/row_probe8.txt | head; grep -c "ok$" gpurun_out/row_probe8.txt; grep -E "^N=.*ctas/SM" gpurun_out/row_probe8.txt | sed -E 's/.*(smem=[0-9]+ regs=[0-9]+ ctas\/SM=[0-9]+\(max [0-9]+\)).*/\1/' | tail -10; grep RESULT gpurun_out/row_probe8.txt | tail -10 | sed -E 's/  RESULT //; s/ dbg=0//; s/TFLOP.*//'
