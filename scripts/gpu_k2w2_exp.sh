#!/bin/bash
# round 2: K2w2 with the converged MMA issuer: parity on the release build, then knob experiments
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "weighted_tensor or smoke" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k2w2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_k2w2.log
timeout 120 python scripts/time_search.py --weighted --tag release_w | tail -1
timeout 120 python scripts/time_search.py --weighted --metric MSE --tag release_w_mse | tail -1
timeout 120 python scripts/time_search.py --tag release_k2 | tail -1
SKY_NVCC_DEFS=-DSKY_EXPERIMENTS python -m sky_embeddings_b200.build --force > gpurun_out/build_exp.log 2>&1; echo "build rc=$?"
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 120 python scripts/time_search.py --weighted --tag $tag 2>gpurun_out/exp_$tag.err | tail -1
}
run spin0 SKY_TW2_SPIN=0
run spin1 SKY_TW2_SPIN=1
run st5 SKY_TW2_SPIN=0 SKY_TW2_STAGES=5
run evictfirst SKY_TW2_SPIN=0 SKY_TW2_POLICY=1
run noepi SKY_TW2_SPIN=0 SKY_TW_DEBUG=4
run nosq_noepi SKY_TW2_SPIN=0 SKY_TW_DEBUG=5
