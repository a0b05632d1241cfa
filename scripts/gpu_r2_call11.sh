#!/bin/bash
# round 2, call 11: ncu --set full (source-level) of the halo conv at the RAB shapes
set -u
OUT=gpurun_out/r2c11
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SR_PROFILE=1 SR_ONLY="G.K" timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -o $OUT/ncu_halo python scripts/conv_bench.py > $OUT/ncu_halo.log 2>&1
tail -3 $OUT/ncu_halo.log
ls -la $OUT
