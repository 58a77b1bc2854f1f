#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest batch"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x -k "batch" > gpurun_out/pytest_batch.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_batch.log
for dbg in ${DBGS:-0 4 1}; do
SKY_TC_DEBUG=$dbg SKY_TB_DEBUG=$dbg timeout 900 python bench.py --workload ${WL:-c3} --steps 3 --warmup 3 --no-cpu ${EXTRA:-} > gpurun_out/bd.json 2> gpurun_out/bd.err; echo "dbg=$dbg rc=$?"; tail -2 gpurun_out/bd.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bd.json")); r=d["roofline"]
    print("dbg=$dbg kernel_ms=%.3f %s=%.1f frac=%.3f step_ms=%.3f clocks=%s" % (r["kernel_ms"], r["unit"], r["achieved"], r["frac"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e: print("ERR", e)
PY
done
