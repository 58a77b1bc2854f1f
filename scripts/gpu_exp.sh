#!/bin/bash
# tests + bench + debug-mode experiments (no ncu)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for dbg in 0 1 2 3; do
  echo "== SKY_TC_DEBUG=$dbg"
  SKY_TC_DEBUG=$dbg timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu ${BENCH_ARGS:-} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    r=d['roofline']; print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']))
"
done
