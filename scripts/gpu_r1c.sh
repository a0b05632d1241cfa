#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== conv tests"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -x --timeout 120 2>&1 | tail -15 | tee gpurun_out/t_conv.log
echo "== fused tests"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_fused_kernels.py -m gpu -q --timeout 300 2>&1 | tail -25 | tee gpurun_out/t_fused.log
echo "== model tests"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_model_parity.py -m gpu -q --timeout 300 2>&1 | tail -25 | tee gpurun_out/t_model.log
echo "== phases"; timeout -s KILL 300 python -W ignore scripts/phase_times.py 2>&1 | tail -14 | tee gpurun_out/phases.log
echo "== bench"; timeout -s KILL 900 python -W ignore bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== torch prof"; timeout -s KILL 400 python -W ignore scripts/torch_prof.py > gpurun_out/torch_prof.log 2>&1
