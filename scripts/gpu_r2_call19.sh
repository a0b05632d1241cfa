#!/bin/bash
# round 2, call 19 (8 GPUs): chain training x2 -> x3 -> x4 at global batch 128 (BASELINE.json configs[2]) + the 8-GPU bench line
set -u
OUT=gpurun_out/r2c19
mkdir -p $OUT
export PYTHONUNBUFFERED=1
N=${NGPU:-8}
TAILN=12 bash scripts/run_ranks.sh $N 420 scripts/chain_bench.py 2>&1 | tee $OUT/summary.txt
cp gpurun_out/rank0.log $OUT/chain_rank0.log
bash scripts/run_torchrun_guarded.sh $N 300 --steps 16 --no-edsr --no-inference --no-comparator --no-cpu-baseline 2>&1 | tee -a $OUT/summary.txt
cp gpurun_out/torchrun_$N.log $OUT/
