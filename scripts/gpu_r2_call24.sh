#!/bin/bash
# round 2, call 24: main chain captured on a high-priority stream vs default
set -u
OUT=gpurun_out/r2c24
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for PR in 1 0 1 0; do
  SR_GRAPH_PRIORITY=$PR timeout -s KILL 200 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 24 > $OUT/bench_pr$PR.json 2> $OUT/bench.err
  python - <<PY | tee -a $OUT/summary.txt
import json
d = json.loads(open("$OUT/bench_pr$PR.json").read().strip().splitlines()[-1])
print("SR_GRAPH_PRIORITY=$PR ms/step %.3f img/s %.1f e2e %.1f loss_finite %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["loss_finite"]))
PY
done
