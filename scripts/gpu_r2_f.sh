#!/bin/bash
# round 2, session 3: K2b with eight epilogue warps -- parity, timing, timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "batch_path or config3 or config4" --timeout 200 --timeout-method=thread -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/f_pytest.log | cut -c1-200
timeout 300 python scripts/ab_knobs.py c3g8r c4g8r c3r 2>&1 | cut -c1-12,66-260
cp sky_embeddings_b200/libskysearch.so /tmp/rel.so
cp sky_embeddings_b200/libskysearch_exp.so sky_embeddings_b200/libskysearch.so
for ph in 1 4; do timeout 120 python tools/trace_tb_phase.py $ph > gpurun_out/f_trace_c3g8_p$ph.txt 2>&1; echo "trace $ph rc=$?"; done
cp /tmp/rel.so sky_embeddings_b200/libskysearch.so
sed -n 1,8p gpurun_out/f_trace_c3g8_p1.txt | cut -c1-200; sed -n 20,26p gpurun_out/f_trace_c3g8_p4.txt | cut -c1-200
echo "t=$(( $(date +%s) - T0 ))"
