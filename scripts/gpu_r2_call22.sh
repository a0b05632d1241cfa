#!/bin/bash
set -u
OUT=gpurun_out/r2c22
mkdir -p $OUT
timeout -s KILL 300 python - <<'PY' 2>&1 | tee $OUT/summary.txt
import sys
sys.path.insert(0, ".")
import bench
for tb in (32, 48, 64, 96):
    r = bench.inference_bench(2048, tile_batch=tb)
    print("tile_batch %3d: %.1f Mpix/s (%.3f s)" % (tb, r["value"], r["seconds"]), flush=True)
PY
