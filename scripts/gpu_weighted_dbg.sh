#!/bin/bash
set -u
mkdir -p gpurun_out
for dbg in ${DBGS:-0 1 2 3 4 7}; do
SKY_TC_DEBUG=1 SKY_TW_DEBUG=$dbg timeout 300 python bench.py --workload c2 --weighted --steps 100 --warmup 5 --no-cpu > gpurun_out/bw.json 2> gpurun_out/bw.err; tail -3 gpurun_out/bw.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bw.json")); r=d["roofline"]
    print("dbg=$dbg kernel_ms=%.4f frac=%.3f step_ms=%.4f" % (r["kernel_ms"], r["frac"], d["ms_per_step"]))
except Exception as e: print("ERR", e)
PY
done
