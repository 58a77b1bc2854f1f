#!/bin/bash
# 2-GPU check of the paths this session touched: fused sharded search (pack -> K2 -> shard merge into the peers, PDL) and
# the batched kernel behind the push / merge exchange; strong-scaled C3 as sub-record
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 200 --warmup 10 --also c3 > gpurun_out/n2_default.json 2> gpurun_out/n2_default.err
echo "N=$N rc=$?"; grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$\|NCCL version\|W[0-9]* " gpurun_out/n2_default.err | tail -5
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/n2_default.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("  c2 value=%.0f global_qps=%.0f e2e=%.0f step_ms=%.4f kernel_ms=%.4f frac=%.3f exchange=%s merge_bit_exact=%s launches/step=%s kernel_ms_per_rank=%s" % (d["value"], d["qps_global_bank"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d.get("exchange"), d["parity"].get("merge_bit_exact"), d.get("gpu_launches_per_step"), d.get("kernel_ms_per_rank")))
    for n, a in d.get("also", {}).items():
        if "skipped" in a: print("   also", n, a); continue
        ar = a["roofline"]
        print("   also %s value=%.0f step_ms=%.3f e2e=%.0f kernel_ms=%.3f frac=%.3f step_frac=%s merge_bit_exact=%s clocks=%s" % (n, a["value"], a["ms_per_step"], a["e2e"]["value"], ar["kernel_ms"], ar["frac"], ar.get("step_frac"), a["parity"].get("merge_bit_exact"), a["clocks"]["sm_mhz"]))
except Exception as e: print("ERR", e)
PY
