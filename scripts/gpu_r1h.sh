#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== conv bench"; timeout -s KILL 300 python -W ignore scripts/conv_bench.py 2>&1 | tail -20 | tee gpurun_out/conv_bench.log
echo "== ncu launch list"; timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python -W ignore scripts/profile_step.py 2>&1 | tail -2
echo "== ncu full"; SR_ONLY="G.K" SR_PROFILE=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|conv_tc_wgrad" -c 6 -f -o gpurun_out/prof_final python -W ignore scripts/conv_bench.py 2>&1 | tail -2
