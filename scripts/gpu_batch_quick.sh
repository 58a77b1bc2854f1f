#!/bin/bash
set -u
mkdir -p gpurun_out
for wl in ${WLS:-c3g8 c4 c3 c3s}; do
timeout 900 python bench.py --workload $wl --steps ${STEPS:-5} --warmup 3 --no-cpu > gpurun_out/bd.json 2> gpurun_out/bd.err; tail -2 gpurun_out/bd.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bd.json")); r=d["roofline"]
    print("$wl kernel_ms=%.3f frac=%.3f step_ms=%.3f qps=%.0f launches=%d clocks=%s" % (r["kernel_ms"], r["frac"], d["ms_per_step"], d["value"], r["kernel_launches"], d["clocks"]["sm_mhz"]))
except Exception as e: print("ERR", e)
PY
done
