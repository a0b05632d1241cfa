#!/bin/bash
set -u
OUT=gpurun_out/call12
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for px in 256 128 96; do
  echo "SR_LA_SLICE_PX=$px" | tee -a $OUT/summary.txt
  SR_LA_SLICE_PX=$px timeout -s KILL 300 python scripts/la_bench.py 2>&1 | grep la_chain | tee -a $OUT/summary.txt
done
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench.json 2> $OUT/bench.err
SR_LA_SLICE_PX=256 timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench_256.json 2> $OUT/bench_256.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for f in ("bench.json", "bench_256.json"):
    try:
        d = json.loads(open("gpurun_out/call12/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $OUT/bench.err
