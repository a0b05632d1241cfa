#!/bin/bash
set -u
OUT=gpurun_out/call11
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/la_bench.py 2>&1 | grep la_chain | tee -a $OUT/summary.txt
SR_LA_SIDE=0 timeout -s KILL 300 python scripts/la_bench.py 2>&1 | grep la_chain | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench.json 2> $OUT/bench.err
SR_LA_SIDE=0 timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench_noside.json 2> $OUT/bench_noside.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for f in ("bench.json", "bench_noside.json"):
    try:
        d = json.loads(open("gpurun_out/call11/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $OUT/bench.err
