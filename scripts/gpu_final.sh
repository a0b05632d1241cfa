#!/bin/bash
# final validation of the round: full GPU suite, smoke, the default bench (all sub-lines), reference arm, launch list
set -u
OUT=gpurun_out/final
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/final/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["step_frac_of_tensor_roofline"])
print("edsr", d["edsr"]); print("inference", d["inference"]["value"]); print("cpu", d["cpu_baseline"])
r = json.loads(open("gpurun_out/final/bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["sample"])
PY
timeout -s KILL 300 python scripts/conv_bench.py > $OUT/conv_bench.txt 2>&1
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 60 > $OUT/launches_summary.txt 2>&1
timeout -s KILL 300 python scripts/infer_sweep.py > $OUT/infer_sweep.txt 2>&1; grep tile $OUT/infer_sweep.txt | tee -a $OUT/summary.txt
