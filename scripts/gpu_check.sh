#!/bin/bash
# One gpurun call: smoke, GPU parity tests, short bench.  Every step is bounded by `timeout`.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; cat gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
