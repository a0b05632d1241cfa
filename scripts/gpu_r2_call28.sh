#!/bin/bash
# round 2, call 28: one multi-tensor launch per critic pass for the BatchNorm counters: tests + step time
set -u
OUT=gpurun_out/r2c28
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_model_parity.py tests/test_gpu_ndsrgan.py -m gpu -q --timeout 300 -x > $OUT/pytest.log 2>&1
echo "pytest exit $?" | tee $OUT/summary.txt
tail -4 $OUT/pytest.log | tee -a $OUT/summary.txt
timeout -s KILL 200 python bench.py --no-edsr --no-inference --no-comparator --no-cpu-baseline --steps 24 > $OUT/bench.json 2> $OUT/bench.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
d = json.loads(open("gpurun_out/r2c28/bench.json").read().strip().splitlines()[-1])
print("ms/step %.3f img/s %.1f e2e %.1f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
PY
