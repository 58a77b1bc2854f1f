#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest weighted"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x -k "weighted" > gpurun_out/pytest_w.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_w.log
for wl in c2 c2mse; do
timeout 300 python bench.py --workload $wl --weighted --steps 100 --warmup 5 --no-cpu > gpurun_out/bw.json 2> gpurun_out/bw.err; echo "$wl weighted rc=$?"; tail -3 gpurun_out/bw.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bw.json")); r=d["roofline"]
    print("  value=%.0f kernel_ms=%.4f launches=%d frac=%.3f step_ms=%.4f" % (d["value"], r["kernel_ms"], r["kernel_launches"], r["frac"], d["ms_per_step"]))
except Exception as e: print("ERR", e)
PY
done
