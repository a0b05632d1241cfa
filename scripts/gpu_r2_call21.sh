#!/bin/bash
set -u
OUT=gpurun_out/r2c21
mkdir -p $OUT
for PDL in 0 1 0 1; do SR_PDL=$PDL timeout -s KILL 120 python scripts/graph_gap_probe.py 2>&1 | tail -1 | tee -a $OUT/summary.txt; done
