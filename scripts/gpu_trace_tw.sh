#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SKY_NVCC_DEFS=-DSKY_EXPERIMENTS python -m sky_embeddings_b200.build --force > gpurun_out/build_exp.log 2>&1; echo "build rc=$?"
SKY_TW_DEBUG=32 timeout 200 python tools/trace_tw.py > gpurun_out/trace_tw.txt 2>&1; echo "trace rc=$?"; tail -3 gpurun_out/trace_tw.txt
SKY_TW_DEBUG=36 timeout 200 python tools/trace_tw.py > gpurun_out/trace_tw_noepi.txt 2>&1; echo "trace rc=$?"; tail -3 gpurun_out/trace_tw_noepi.txt
