#!/bin/bash
# round 2, call 14: resample kernel vs PIL, DP overlap check on one GPU (four-segment replay), full GPU suite
set -u
OUT=gpurun_out/r2c14
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_input_pipeline.py -m gpu -q --timeout 300 > $OUT/pytest_input.log 2>&1
echo "pytest(input) exit $?" | tee $OUT/summary.txt
tail -5 $OUT/pytest_input.log | tee -a $OUT/summary.txt
timeout -s KILL 600 python scripts/dp_overlap_check.py > $OUT/dp_overlap_1gpu.txt 2>&1
tail -5 $OUT/dp_overlap_1gpu.txt | tee -a $OUT/summary.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x --deselect tests/test_gpu_input_pipeline.py > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -6 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
