#!/bin/bash
set -u
OUT=gpurun_out/call6
mkdir -p $OUT
export PYTHONUNBUFFERED=1
python scripts/diag_grad_err.py > $OUT/grad_err_mma.txt 2>&1
grep -v Warn $OUT/grad_err_mma.txt | head -8 | tee $OUT/summary.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest (async wgrad) exit $?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
SR_WGRAD_ASYNC=0 timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_gpu_sync.log 2>&1
echo "pytest (sync wgrad) exit $?" | tee -a $OUT/summary.txt
tail -4 $OUT/pytest_gpu_sync.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/la_bench.py 2>&1 | grep la_chain | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference > $OUT/bench.json 2> $OUT/bench.err
SR_WGRAD_ASYNC=0 timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference > $OUT/bench_sync.json 2> $OUT/bench_sync.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for f in ("bench.json", "bench_sync.json"):
    try:
        d = json.loads(open("gpurun_out/call6/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "edsr", d["edsr"].get("ms_per_step"), d["edsr"].get("tflops"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $OUT/bench.err
