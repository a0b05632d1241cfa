#!/bin/bash
# session-3 GPU call 4: mma.sync LA kernels + row-wise reduce: kernel tests first, then suite + bench (with/without the new LA path)
set -u
OUT=gpurun_out/call4
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_fused_kernels.py tests/test_gpu_conv_kernels.py -q --timeout 120 -k "la_chain or wgrad" > $OUT/kernels.log 2>&1
echo "kernel tests exit $?" | tee $OUT/summary.txt
tail -15 $OUT/kernels.log | tee -a $OUT/summary.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -8 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/la_bench.py > $OUT/la_bench_mma.txt 2>&1
SR_LA_MMA=0 timeout -s KILL 300 python scripts/la_bench.py > $OUT/la_bench_simt.txt 2>&1
tail -4 $OUT/la_bench_mma.txt $OUT/la_bench_simt.txt | tee -a $OUT/summary.txt
timeout -s KILL 300 python scripts/conv_bench.py > $OUT/conv_bench.txt 2>&1
tail -3 $OUT/conv_bench.txt | tee -a $OUT/summary.txt
timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench.json 2> $OUT/bench.err
SR_LA_MMA=0 timeout -s KILL 900 python bench.py --no-cpu-baseline --no-inference --no-edsr > $OUT/bench_simt_la.json 2> $OUT/bench_simt_la.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for f in ("bench.json", "bench_simt_la.json"):
    try:
        d = json.loads(open("gpurun_out/call4/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
PY
