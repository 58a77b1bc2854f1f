#!/bin/bash
# One gpurun --gpus N call: torchrun bench lines at N ranks; the candidate exchange over peer memory vs NCCL.
# The bench gates every line on a bit-exact comparison of the exchanged + merged top-k with a torch merge of the
# independently all-gathered candidates (parity.merge_bit_exact).
set -u
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
run() { # name, args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/multi_${name}_n$N.json 2> gpurun_out/multi_${name}_n$N.err
  echo "$name N=$N rc=$?"; grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$\|NCCL version\|W[0-9]* " gpurun_out/multi_${name}_n$N.err | tail -4
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/multi_${name}_n$N.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("  value=%.0f global_qps=%.0f e2e=%.0f step_ms=%.4f kernel_ms=%.4f %s frac=%.3f exchange=%s merge_bit_exact=%s launches/step=%s" % (d["value"], d["qps_global_bank"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["unit"], r["frac"], d.get("exchange"), d["parity"].get("merge_bit_exact"), d.get("gpu_launches_per_step")))
    for n, a in d.get("also", {}).items():
        if "skipped" in a: print("   also", n, a); continue
        ar = a["roofline"]
        print("   also %s value=%.0f step_ms=%.3f kernel_ms=%.3f frac=%.3f step_frac=%s merge_bit_exact=%s clocks=%s" % (n, a["value"], a["ms_per_step"], ar["kernel_ms"], ar["frac"], ar.get("step_frac"), a["parity"].get("merge_bit_exact"), a["clocks"]["sm_mhz"]))
except Exception as e: print("ERR", e)
PY
}
run c2_peer --steps 200 --warmup 10 --no-cpu --also none --exchange peer
run c2_nccl --steps 200 --warmup 10 --no-cpu --also none --exchange nccl
[ "$N" -le 2 ] && run c2w_peer --workload c2w --steps 100 --warmup 10 --no-cpu --also none --exchange peer
if [ "${FULL:-0}" = "1" ]; then
  run default_peer --no-cpu --exchange peer
fi
