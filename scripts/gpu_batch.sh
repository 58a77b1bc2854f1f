#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest batch"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x -k "batch" > gpurun_out/pytest_batch.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_batch.log
for wl in ${WLS:-c3s c3}; do
echo "== bench $wl"; timeout 900 python bench.py --workload $wl --steps ${STEPS:-5} --warmup 3 --no-cpu ${EXTRA:-} > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "rc=$?"; tail -3 gpurun_out/bench_$wl.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$wl.json")); r=d["roofline"]
    print("$wl value=%.1f e2e=%.1f kernel_ms=%.3f %s=%.1f frac=%.3f step_ms=%.3f clocks=%s" % (d["value"], d["e2e"]["value"], r["kernel_ms"], r["unit"], r["achieved"], r["frac"], d["ms_per_step"], d["clocks"]))
except Exception as e: print("ERR", e)
PY
done
