#!/bin/bash
# round-2 evidence: launch list of the bench command and one --set full capture of the split-FP16 tensor kernel
set -x
CMD="python bench.py --steps 5 --warmup 3 --no-sweep --no-configs --sustain-seconds 0 --cpu-seconds 1 --e2e-steps 2"
$CMD > gpurun_out/r02_plain.log 2> gpurun_out/r02_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fir_tc2 -s 3 -c 2 -o gpurun_out/r02_tc2 $CMD > gpurun_out/r02_ncu2.log 2>&1
tail -3 gpurun_out/r02_ncu2.log
