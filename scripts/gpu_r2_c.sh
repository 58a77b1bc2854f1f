#!/bin/bash
# round 2, session 3: K2b survivor queue -- parity (batch path, full-size C3 / C4), then old library vs new, alternating
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q -x -k "batch or tensor_path or sharded or search_host" --timeout 200 --timeout-method=thread -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/c_pytest.log | cut -c1-200
timeout -k 10 400 python -m pytest tests/test_gpu_fullsize.py -q -x -k "config3 or config4" --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/c_pytest_full.log 2>&1; echo "pytest full rc=$? t=$(( $(date +%s) - T0 ))"; tail -3 gpurun_out/c_pytest_full.log | cut -c1-200
cp sky_embeddings_b200/libskysearch.so /tmp/new.so
for round in 1 2; do
  for which in old new; do
    if [ $which = old ]; then cp sky_embeddings_b200/libskysearch_old.so sky_embeddings_b200/libskysearch.so; else cp /tmp/new.so sky_embeddings_b200/libskysearch.so; fi
    echo "== $which (round $round)"
    timeout 300 python scripts/ab_knobs.py c3g8r c4g8r c2r 2>&1 | cut -c1-40,101-260
  done
done
cp /tmp/new.so sky_embeddings_b200/libskysearch.so
echo "== new: c3 full"; timeout 200 python scripts/ab_knobs.py c3r 2>&1 | cut -c1-40,101-260
echo "t=$(( $(date +%s) - T0 ))"
