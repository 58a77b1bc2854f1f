#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "batch or tensor_path or config2 or sharded" 2>&1 | tail -3
for wl in mid c3s; do for path in auto tensor batch; do
timeout 300 python bench.py --workload $wl --path $path --steps 20 --warmup 5 --no-cpu > gpurun_out/bd.json 2> gpurun_out/bd.err; tail -2 gpurun_out/bd.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bd.json")); r=d["roofline"]
    print("$wl $path step_ms=%.4f qps=%.0f launches/step=%.1f" % (d["ms_per_step"], d["value"], d["gpu_launches"]/d["steps"]))
except Exception as e: print("ERR", e)
PY
done; done
