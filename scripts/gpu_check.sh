#!/bin/bash
# Runs on the GPU box under gpurun: kernel tests, model parity, smoke, short bench. Logs -> gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$1" != "quick" ]; then
echo "== simt kernels"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -k "simt or thin" -x --timeout 300 2>&1 | tail -15 | tee gpurun_out/t_simt.log
echo "== tcgen05 kernels"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -k "tcgen05 or full_size" --timeout 120 2>&1 | tail -40 | tee gpurun_out/t_tc.log
fi
echo "== fused kernels"; timeout -s KILL 600 python -W ignore -m pytest tests/test_gpu_fused_kernels.py -m gpu -q --timeout 120 2>&1 | tail -30 | tee gpurun_out/t_fused.log
echo "== model parity"; timeout -s KILL 900 python -W ignore -m pytest tests/test_gpu_model_parity.py -m gpu -q --timeout 300 2>&1 | tail -40 | tee gpurun_out/t_model.log
echo "== smoke"; timeout -s KILL 300 python -W ignore __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"; timeout -s KILL 900 python -W ignore bench.py --steps 4 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
if [ "$2" == "ncu" ]; then
echo "== ncu launch list"; timeout -s KILL 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python -W ignore scripts/profile_step.py 2>&1 | tail -3
fi
