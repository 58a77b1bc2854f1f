#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest batch"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x -k "batch" > gpurun_out/pytest_batch.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_batch.log
for gr in 2 4 8; do for wl in c3g8 c4 c3; do
SKY_TB_GROWTH=$gr timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/bd.json 2> gpurun_out/bd.err; tail -2 gpurun_out/bd.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bd.json")); r=d["roofline"]
    print("growth=$gr $wl kernel_ms=%.3f frac=%.3f step_ms=%.3f launches=%d clocks=%s" % (r["kernel_ms"], r["frac"], d["ms_per_step"], r["kernel_launches"], d["clocks"]["sm_mhz"]))
except Exception as e: print("ERR", e)
PY
done; done
