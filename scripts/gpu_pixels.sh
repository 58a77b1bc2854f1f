#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest pixel"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x -k "pixel" > gpurun_out/pytest_pixel.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_pixel.log
for wl in ${WLS:-c5s c5q1 c5}; do
echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "rc=$?"; tail -3 gpurun_out/bench_$wl.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$wl.json")); r=d["roofline"]
    print("$wl value=%.1f e2e=%.1f kernel_ms=%.3f GB/s=%.0f frac=%.3f step_ms=%.3f cpu=%s" % (d["value"], d["e2e"]["value"], r["kernel_ms"], r["achieved"], r["frac"], d["ms_per_step"], d.get("cpu_baseline",{}).get("value")))
except Exception as e: print("ERR", e)
PY
done
