#!/bin/bash
# round 2, call 18: SRGAN sibling on the GPU
set -u
OUT=gpurun_out/r2c18
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests/test_gpu_srgan.py -m gpu -q --timeout 600 > $OUT/pytest_srgan.log 2>&1
echo "pytest(srgan) exit $?" | tee $OUT/summary.txt
tail -40 $OUT/pytest_srgan.log | tee -a $OUT/summary.txt
