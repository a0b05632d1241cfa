#!/bin/bash
set -u
OUT=gpurun_out/call8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -6 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for f in ("bench.json",):
    try:
        d = json.loads(open("gpurun_out/call8/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "edsr", d["edsr"].get("ms_per_step"), d["edsr"].get("tflops"))
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]): print("  ", k, round(v["ms_per_step"], 3), v["launches_per_step"], round(v["tflops"] or 0, 1))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $OUT/bench.err
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 40 > $OUT/launches_summary.txt 2>&1
head -30 $OUT/launches_summary.txt | tee -a $OUT/summary.txt
