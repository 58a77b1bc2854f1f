#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
b() { name=$1; shift
timeout 900 python bench.py "$@" --no-cpu > gpurun_out/bd.json 2> gpurun_out/bd.err; tail -2 gpurun_out/bd.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bd.json")); r=d["roofline"]
    print("$name kernel_ms=%.4f frac=%.3f step_ms=%.4f qps=%.0f e2e=%.0f" % (r["kernel_ms"], r["frac"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e: print("ERR", e)
PY
}
b c2 --steps 200 --warmup 10
b q1_bf16 --workload q1 --steps 50 --warmup 5
b q1_fp32_w --workload q1 --bank-dtype fp32 --weighted --steps 50 --warmup 5
b c2_weighted --weighted --steps 100 --warmup 5
b c3g8 --workload c3g8 --steps 5 --warmup 3
b c4 --workload c4 --steps 5 --warmup 3
b c3 --workload c3 --steps 5 --warmup 3
