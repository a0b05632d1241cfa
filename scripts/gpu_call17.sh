#!/bin/bash
set -u
OUT=gpurun_out/call17
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
for rep in 1 2; do
  timeout -s KILL 600 python bench.py --no-cpu-baseline --no-inference --no-edsr --steps 16 > $OUT/b$rep.json 2> $OUT/b.err
  python - $rep <<'PY' | tee -a $OUT/summary.txt
import json, sys
d = json.loads(open("gpurun_out/call17/b%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.3f" % d["ms_per_step"], "img/s %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"])
PY
done
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py > $OUT/ncu_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches.csv 60 > $OUT/launches_summary.txt 2>&1
grep -E "^launches|bn_|colsum" $OUT/launches_summary.txt | tee -a $OUT/summary.txt
