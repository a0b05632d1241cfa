#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== conv tests"; timeout -s KILL 400 python -W ignore -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -x --timeout 60 -k "wgrad or full_size" 2>&1 | tail -8
echo "== conv bench"; timeout -s KILL 300 python -W ignore scripts/conv_bench.py 2>&1 | tail -20 | tee gpurun_out/conv_bench.log
