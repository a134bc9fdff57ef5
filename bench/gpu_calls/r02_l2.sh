#!/bin/bash
{
for t in 51 251 1001; do
  for env in "" "NOL2=1"; do echo -n "$env 16blk: "; env $env timeout 30 bench/microbench/tc2_probe $t 16384 512 16 5 1 | tail -1 | cut -c1-330; done
  for env in "" "NOL2=1"; do echo -n "$env 1blk: "; env $env timeout 30 bench/microbench/tc2_probe $t 65536 1024 1 5 1 | tail -1 | cut -c1-330; done
done
for u in 2 4; do for env in "" "NOL2=1"; do echo -n "$env U=$u 1blk: "; env $env timeout 30 bench/microbench/tc2_probe 251 65536 1024 1 5 $u | tail -1 | cut -c1-200; done; done
} > gpurun_out/tc2_l2_probe.txt 2>&1
cut -c1-260 gpurun_out/tc2_l2_probe.txt
