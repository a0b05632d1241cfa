#!/bin/bash
# session-3 GPU call 3: full suite (incl. EDSR), bench, ncu launch list of one step, ncu --set full of the RAB wgrad launches
set -u
OUT=gpurun_out/call3
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/conv_bench.py > $OUT/conv_bench.txt 2>&1
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 70 > $OUT/launches_summary.txt 2>&1
SR_PROFILE=1 SR_ONLY=G.K timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:wgrad -o $OUT/wgrad_full python scripts/conv_bench.py > $OUT/ncu_wgrad.log 2>&1
ncu -i $OUT/wgrad_full.ncu-rep --page raw --csv > $OUT/wgrad_full_raw.csv 2>/dev/null
ls -la $OUT | tee -a $OUT/summary.txt
