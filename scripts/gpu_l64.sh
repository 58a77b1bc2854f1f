#!/bin/bash
set -u
mkdir -p gpurun_out
for dt in fp32 bf16; do for wf in "" "--weighted"; do
timeout 300 python bench.py --workload l64 --bank-dtype $dt $wf --steps 30 --warmup 5 --no-cpu > gpurun_out/l64.json 2> gpurun_out/l64.err; echo "l64 $dt $wf rc=$?"; tail -2 gpurun_out/l64.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/l64.json")); r=d["roofline"]
    print("  kernel_ms=%.4f %s=%.0f frac=%.3f step_ms=%.4f" % (r["kernel_ms"], r["unit"], r["achieved"], r["frac"], d["ms_per_step"]))
except Exception as e: print("ERR", e)
PY
done; done
