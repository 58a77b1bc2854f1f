#!/bin/bash
set -u
mkdir -p gpurun_out
for st in 6 8 10 12; do for cfg in "q1 fp32 --weighted" "q1 bf16" "l64 fp32 --weighted"; do
set -- $cfg
SKY_ST_STAGES=$st timeout 200 python bench.py --workload $1 --bank-dtype $2 ${3:-} --steps 40 --warmup 5 --no-cpu > gpurun_out/sw.json 2> gpurun_out/sw.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw.json")); r=d["roofline"]
    print("stages=$st $cfg kernel_ms=%.4f GB/s=%.0f step_ms=%.4f" % (r["kernel_ms"], r["achieved"], d["ms_per_step"]))
except Exception as e: print("ERR", e, open("gpurun_out/sw.err").read()[-200:])
PY
done; done
