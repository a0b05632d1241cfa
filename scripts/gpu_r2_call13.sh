#!/bin/bash
# round 2, call 13 (2 GPUs): DP overlap check + 2-GPU bench; cbam test re-run
set -u
OUT=gpurun_out/r2c13
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_cbam_kernels.py -m gpu -q --timeout 300 -x > $OUT/pytest_cbam.log 2>&1
echo "pytest(cbam) exit $?" | tee $OUT/summary.txt
tail -5 $OUT/pytest_cbam.log | tee -a $OUT/summary.txt
TAILN=8 bash scripts/run_ranks.sh 2 400 scripts/dp_overlap_check.py 2>&1 | tee -a $OUT/summary.txt
cp gpurun_out/rank0.log $OUT/dp_overlap_rank0.log
bash scripts/run_torchrun_guarded.sh 2 400 --steps 16 --no-edsr --no-inference --no-comparator --no-cpu-baseline 2>&1 | tee -a $OUT/summary.txt
cp gpurun_out/torchrun_2.log $OUT/
