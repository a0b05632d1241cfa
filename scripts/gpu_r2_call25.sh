#!/bin/bash
# round 2, call 25: SRAGAN / SRGAN siblings on the GPU
set -u
OUT=gpurun_out/r2c25
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 500 python -m pytest tests/test_gpu_sragan.py tests/test_gpu_srgan.py -m gpu -q --timeout 300 > $OUT/pytest_siblings.log 2>&1
echo "pytest(siblings) exit $?" | tee $OUT/summary.txt
tail -40 $OUT/pytest_siblings.log | tee -a $OUT/summary.txt
