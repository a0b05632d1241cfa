#!/bin/bash
# first measurement batch of this session: tests, bench, conv microbench, ncu launch list, ncu full of the top kernel
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu"; timeout -s KILL 900 python -W ignore -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== diag"; timeout -s KILL 300 python -W ignore scripts/diag_dstate.py 2>&1 | tail -80 | tee gpurun_out/diag.log
echo "== conv bench"; timeout -s KILL 300 python -W ignore scripts/conv_bench.py 2>&1 | tail -30 | tee gpurun_out/conv_bench.log
echo "== phases"; timeout -s KILL 300 python -W ignore scripts/phase_times.py 2>&1 | tail -30 | tee gpurun_out/phases.log
echo "== bench"; timeout -s KILL 900 python -W ignore bench.py --steps 8 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== ncu launch list"; timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python -W ignore scripts/profile_step.py 2>&1 | tail -3
echo "== ncu full K1/K2"; SR_ONLY="G.K" SR_PROFILE=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 6 -f -o gpurun_out/prof_conv python -W ignore scripts/conv_bench.py 2>&1 | tail -5
