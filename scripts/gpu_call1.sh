#!/bin/bash
# Round-1 session-3 GPU call 1: validate the halo weight-gradient kernel + batched packing, then the whole suite and the bench.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/call1
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > $OUT/gpu.txt 2>&1

echo "== wgrad parity, halo kernel ==" | tee $OUT/summary.txt
SR_WG_HALO=1 timeout -s KILL 240 python -m pytest tests/test_gpu_conv_kernels.py -q --timeout 60 -k "wgrad or batched" > $OUT/wgrad_halo.log 2>&1
H=$?
tail -3 $OUT/wgrad_halo.log | tee -a $OUT/summary.txt
if [ $H -ne 0 ]; then
  echo "halo wgrad FAILED parity -> falling back to SR_WG_HALO=0 for the rest" | tee -a $OUT/summary.txt
  export SR_WG_HALO=0
else
  export SR_WG_HALO=1
fi

echo "== conv bench (wgrad column), old kernel ==" | tee -a $OUT/summary.txt
SR_WG_HALO=0 timeout 300 python scripts/conv_bench.py > $OUT/conv_bench_old.txt 2>&1
tail -20 $OUT/conv_bench_old.txt | tee -a $OUT/summary.txt
if [ $H -eq 0 ]; then
  echo "== conv bench, halo wgrad ==" | tee -a $OUT/summary.txt
  SR_WG_HALO=1 timeout 300 python scripts/conv_bench.py > $OUT/conv_bench_halo.txt 2>&1
  tail -20 $OUT/conv_bench_halo.txt | tee -a $OUT/summary.txt
  echo "== conv bench, halo wgrad, RMULT=2 ==" | tee -a $OUT/summary.txt
  SR_WG_HALO=1 SR_WG_RMULT=2 timeout 300 python scripts/conv_bench.py > $OUT/conv_bench_halo_r2.txt 2>&1
  tail -20 $OUT/conv_bench_halo_r2.txt | tee -a $OUT/summary.txt
fi

echo "== full gpu test suite (SR_WG_HALO=$SR_WG_HALO) ==" | tee -a $OUT/summary.txt
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/summary.txt
tail -5 $OUT/pytest_gpu.log | tee -a $OUT/summary.txt

echo "== bench ==" | tee -a $OUT/summary.txt
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?" | tee -a $OUT/summary.txt
cat $OUT/bench.json | tee -a $OUT/summary.txt
SR_PACK_PLAN=0 timeout 600 python bench.py --no-inference --no-cpu-baseline > $OUT/bench_noplan.json 2> $OUT/bench_noplan.err
python - <<'PY' | tee -a $OUT/summary.txt
import json
for f in ("bench.json", "bench_noplan.json"):
    try:
        d = json.loads(open("gpurun_out/call1/" + f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
PY
