#!/bin/bash
# round 2, session 3: parity of the touched kernels, pixel bench with float counts, knob A/B (experiments build), launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 420 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -4 gpurun_out/b_pytest.log | cut -c1-200
timeout 400 python bench.py --workload c5 --also c5q1 --no-cpu --steps 20 --warmup 3 > gpurun_out/b_c5.json 2> gpurun_out/b_c5.err; echo "c5 rc=$? t=$(( $(date +%s) - T0 ))"; tail -2 gpurun_out/b_c5.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/b_c5.json').read().strip().splitlines()[-1])
    for n, r in [('c5', d)] + list(d.get('also', {}).items()):
        rf = r['roofline']; print(n, 'ms/step', round(r['ms_per_step'], 3), 'kernel_ms', round(rf['kernel_ms'], 3), 'frac', round(rf['frac'], 3), 'value', round(r['value'], 1))
except Exception as e:
    print('c5 summary failed', e)
PY
cp sky_embeddings_b200/libskysearch.so /tmp/rel.so
cp sky_embeddings_b200/libskysearch_exp.so sky_embeddings_b200/libskysearch.so
timeout 600 python scripts/ab_knobs.py c2 c2w q1w q1wb c3g8 c4g8 c3 > gpurun_out/b_ab.log 2>&1; echo "ab rc=$? t=$(( $(date +%s) - T0 ))"; cat gpurun_out/b_ab.log | cut -c1-260
cp /tmp/rel.so sky_embeddings_b200/libskysearch.so
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/b_c3g8_launches.csv python bench.py --workload c3g8 --also none --no-cpu --steps 2 --warmup 3 > gpurun_out/b_ncu.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - T0 ))"
