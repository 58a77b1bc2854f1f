#!/bin/bash
# round 2: K2w2 experiments (knob build made on the box; the shipped .so is untouched at home)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
# release build first: parity
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "weighted_tensor or smoke or multi_query or tensor_path or merge" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_k2w2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_k2w2.log
timeout 120 python scripts/time_search.py --weighted --tag release_w | tail -1
timeout 120 python scripts/time_search.py --tag release_k2 | tail -1
SKY_NVCC_DEFS=-DSKY_EXPERIMENTS python -m sky_embeddings_b200.build --force > gpurun_out/build_exp.log 2>&1; echo "build rc=$?"
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 120 python scripts/time_search.py --weighted --tag $tag 2>gpurun_out/exp_$tag.err | tail -1
  grep "sky\]" gpurun_out/exp_$tag.err | head -1
}
run spin1 SKY_TW2_SPIN=1
run spin0 SKY_TW2_SPIN=0
run old SKY_TW_PAIR=0
run st6 SKY_TW2_STAGES=6
run st4 SKY_TW2_STAGES=4
run evictfirst SKY_TW2_POLICY=1
run noepi SKY_TW_DEBUG=4
