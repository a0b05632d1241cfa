#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== probe"; timeout -s KILL 120 python -W ignore scripts/umma_probe.py 2>&1 | tail -30 | tee gpurun_out/umma_probe.log
echo "== conv tests"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -x --timeout 120 2>&1 | tail -15 | tee gpurun_out/t_conv.log
echo "== model tests"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_model_parity.py tests/test_gpu_fused_kernels.py -m gpu -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/t_model.log
echo "== conv bench"; timeout -s KILL 300 python -W ignore scripts/conv_bench.py 2>&1 | tail -30 | tee gpurun_out/conv_bench.log
echo "== bench"; timeout -s KILL 900 python -W ignore bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench.log
