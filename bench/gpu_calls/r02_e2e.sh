#!/bin/bash
python bench/microbench/pcie_bw.py
for fmt in i24 i24_packed i16; do for wc in "" "--wc"; do python bench/e2e_probe.py --fmt $fmt $wc 2>&1 | tail -1; done; done
python bench/e2e_probe.py --fmt i24_packed --blocks 32 2>&1 | tail -1
python bench/e2e_probe.py --fmt i24 --blocks 32 2>&1 | tail -1
