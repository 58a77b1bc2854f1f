#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
# release build first: the new tests of this batch
timeout 900 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_cli.py tests/test_gpu_parity.py -m gpu -q -x -k "exchange or cli or pixel" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_new.log
timeout 300 python bench.py --workload c5 --also none --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; tail -2 gpurun_out/bench_c5.err
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1])
    rf = r['roofline']
    print(f"c5 value={r['value']:.1f} ms={r['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} cfg={r['config']['workload'][:80]}")
except Exception as e:
    print('summary failed', e)
PY
SKY_NVCC_DEFS=-DSKY_EXPERIMENTS python -m sky_embeddings_b200.build --force > gpurun_out/build_exp.log 2>&1; echo "build rc=$?"
SKY_TW_DEBUG=32 timeout 200 python tools/trace_tw.py > gpurun_out/trace_tw.txt 2>&1; echo "trace rc=$?"; tail -1 gpurun_out/trace_tw.txt
SKY_TW_DEBUG=36 timeout 200 python tools/trace_tw.py > gpurun_out/trace_tw_noepi.txt 2>&1; echo "trace rc=$?"; tail -1 gpurun_out/trace_tw_noepi.txt
