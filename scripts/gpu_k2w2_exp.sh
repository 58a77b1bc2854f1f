#!/bin/bash
# round 2: K2w2 with incremental descriptors: parity on the release build, timing, bench lines, timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -x -k "weighted_tensor or smoke or cli" --timeout 600 -p no:cacheprovider > gpurun_out/pytest_k2w2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_k2w2.log
timeout 120 python scripts/time_search.py --weighted --tag release_w | tail -1
timeout 120 python scripts/time_search.py --weighted --metric MSE --tag release_w_mse | tail -1
for wl in c2w c2 c5; do
timeout 300 python bench.py --workload $wl --also none --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err
python - $wl <<'PY'
import json, sys
wl = sys.argv[1]
try:
    r = json.loads(open(f'gpurun_out/bench_{wl}.json').read().strip().splitlines()[-1])
    rf = r['roofline']
    print(f"{wl} value={r['value']:.1f} ms={r['ms_per_step']:.4f} e2e_ms={r['e2e']['ms_per_step']:.4f} kern_ms={rf['kernel_ms']:.4f} frac={rf['frac']:.3f} step_frac={rf.get('step_frac')} launches/step={r['gpu_launches_per_step']} parity={r['parity'].get('max_rel_err')}")
except Exception as e:
    print('summary failed', e)
PY
done
SKY_NVCC_DEFS=-DSKY_EXPERIMENTS python -m sky_embeddings_b200.build --force > gpurun_out/build_exp.log 2>&1; echo "build rc=$?"
SKY_TW_DEBUG=32 timeout 200 python tools/trace_tw.py > gpurun_out/trace_tw.txt 2>&1; echo "trace rc=$?"; tail -1 gpurun_out/trace_tw.txt
SKY_TW_DEBUG=36 timeout 200 python tools/trace_tw.py > gpurun_out/trace_tw_noepi.txt 2>&1; echo "trace rc=$?"; tail -1 gpurun_out/trace_tw_noepi.txt
