#!/bin/bash
# round 2: K2w2 with the short fast path: parity on the release build, timing, timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "weighted_tensor or smoke" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k2w2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_k2w2.log
timeout 120 python scripts/time_search.py --weighted --tag release_w | tail -1
timeout 120 python scripts/time_search.py --weighted --metric MSE --tag release_w_mse | tail -1
timeout 120 python scripts/time_search.py --weighted --q 128 --tag release_w_q128 | tail -1
timeout 300 python bench.py --workload c2w --also none --no-cpu --steps 50 --warmup 5 > gpurun_out/bench_c2w.json 2> gpurun_out/bench_c2w.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_c2w.err
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_c2w.json').read().strip().splitlines()[-1])
    rf = r['roofline']
    print(f"c2w value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac')} launches/step={r['gpu_launches_per_step']} parity={r['parity'].get('max_rel_err')}")
except Exception as e:
    print('summary failed', e)
PY
timeout 300 python bench.py --workload c2 --also none --no-cpu --steps 200 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_c2.err
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
    rf = r['roofline']
    print(f"c2 value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac')} launches/step={r['gpu_launches_per_step']} parity={r['parity'].get('max_rel_err')}")
except Exception as e:
    print('summary failed', e)
PY
SKY_NVCC_DEFS=-DSKY_EXPERIMENTS python -m sky_embeddings_b200.build --force > gpurun_out/build_exp.log 2>&1; echo "build rc=$?"
SKY_TW_DEBUG=32 timeout 200 python tools/trace_tw.py > gpurun_out/trace_tw.txt 2>&1; echo "trace rc=$?"; tail -1 gpurun_out/trace_tw.txt
