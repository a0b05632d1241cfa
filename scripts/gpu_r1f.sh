#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== conv tests"; timeout -s KILL 400 python -W ignore -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -x --timeout 60 2>&1 | tail -5
echo "== conv bench halo"; SR_IMPL=3 timeout -s KILL 300 python -W ignore scripts/conv_bench.py 2>&1 | tail -20 | tee gpurun_out/conv_bench_halo.log
echo "== conv bench im2col"; SR_IMPL=2 timeout -s KILL 300 python -W ignore scripts/conv_bench.py 2>&1 | tail -20 | tee gpurun_out/conv_bench_im2col.log
