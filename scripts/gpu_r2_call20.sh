#!/bin/bash
# round 2, call 20: per-shape A/B of the quad-transposed stores and the stack mode over ALL convolution shapes of the step
set -u
OUT=gpurun_out/r2c20
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SR_ITERS=30 timeout -s KILL 200 python scripts/conv_bench.py > $OUT/conv_bench_default.txt 2>&1
SR_ITERS=30 SR_HALO_QUADS=0 timeout -s KILL 200 python scripts/conv_bench.py > $OUT/conv_bench_noquads.txt 2>&1
tail -3 $OUT/conv_bench_default.txt $OUT/conv_bench_noquads.txt
