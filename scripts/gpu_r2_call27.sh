#!/bin/bash
set -u
OUT=gpurun_out/r2c27
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 400 python -m pytest tests/test_gpu_ndsrgan.py -m gpu -q --timeout 300 > $OUT/pytest_ndsrgan.log 2>&1
echo "pytest(ndsrgan) exit $?" | tee $OUT/summary.txt
tail -40 $OUT/pytest_ndsrgan.log | tee -a $OUT/summary.txt
