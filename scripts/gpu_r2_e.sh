#!/bin/bash
# round 2, session 3: pixel kernel with the PTX accumulate (8 x 8 and 16 x 4 shapes), per-phase timeline of K2b's epilogue
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "pixel or config5" --timeout 150 --timeout-method=thread -p no:cacheprovider > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/e_pytest.log | cut -c1-200
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    r = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rf = r['roofline']; print(sys.argv[2], 'ms/step', round(r['ms_per_step'], 3), 'kernel_ms', round(rf['kernel_ms'], 3), 'frac', round(rf['frac'], 3), 'value', round(r['value'], 1), r['clocks']['sm_mhz'], r['clocks']['reasons'])
except Exception as e:
    print(sys.argv[2], 'summary failed', e)
PY
}
timeout 200 python bench.py --workload c5 --also none --no-cpu --steps 20 --warmup 3 > gpurun_out/e_c5.json 2> gpurun_out/e_c5.err; echo "c5 rc=$? t=$(( $(date +%s) - T0 ))"; show gpurun_out/e_c5.json "c5 8x8 ptx"
cp sky_embeddings_b200/libskysearch.so /tmp/rel.so
cp sky_embeddings_b200/libskysearch_exp.so sky_embeddings_b200/libskysearch.so
SKY_PX_WIDE=1 timeout 200 python bench.py --workload c5 --also none --no-cpu --steps 20 --warmup 3 --allow-knobs > gpurun_out/e_c5w.json 2> gpurun_out/e_c5w.err; echo "c5 wide rc=$? t=$(( $(date +%s) - T0 ))"; show gpurun_out/e_c5w.json "c5 16x4 ptx"
SKY_PX_WIDE=0 timeout 200 python bench.py --workload c5 --also none --no-cpu --steps 20 --warmup 3 --allow-knobs > gpurun_out/e_c5n.json 2> gpurun_out/e_c5n.err; show gpurun_out/e_c5n.json "c5 8x8 ptx (exp build)"
for ph in 1 2 4; do timeout 120 python tools/trace_tb_phase.py $ph > gpurun_out/e_trace_c3g8_p$ph.txt 2>&1; echo "trace $ph rc=$?"; done
timeout 120 python tools/trace_tb_phase.py 2 12500000 1000 1000 cosine > gpurun_out/e_trace_c4g8_p2.txt 2>&1; echo "trace c4 rc=$?"
cp /tmp/rel.so sky_embeddings_b200/libskysearch.so
head -24 gpurun_out/e_trace_c3g8_p1.txt | cut -c1-200
echo "t=$(( $(date +%s) - T0 ))"
