#!/bin/bash
# round 2: ingest GPU tests + one ncu --set full capture of the CTA-pair weighted kernel (source-level stalls)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ingest.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_ingest.log 2>&1; echo "pytest ingest rc=$?"; tail -15 gpurun_out/pytest_ingest.log
timeout 200 python scripts/time_search.py --weighted --steps 3 --tag plain > gpurun_out/plain.log 2>&1 && tail -1 gpurun_out/plain.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_weighted2 -s 2 -c 1 -o gpurun_out/k2w2 python scripts/time_search.py --weighted --steps 3 --tag ncu > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
ls -la gpurun_out/*.ncu-rep
