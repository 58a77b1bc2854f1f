#!/bin/bash
# GPU parity tests, then bench lines for the streaming (small-Q) regime.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for wl in ${WLS:-q1 q4}; do for dt in bf16 fp32; do
  echo "== stream $wl $dt"
  timeout 300 python bench.py --workload $wl --bank-dtype $dt --path simt --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${wl}_${dt}.json 2> gpurun_out/bench_${wl}_${dt}.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}_${dt}.json"))
    print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as e: print("ERR", e)
PY
  tail -3 gpurun_out/bench_${wl}_${dt}.err
done; done
