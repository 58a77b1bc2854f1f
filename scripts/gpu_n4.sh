#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 100 --warmup 5 --no-cpu > gpurun_out/multi_c2_n4.json 2> gpurun_out/multi_c2_n4.err; echo "rc=$?"
grep "^{" gpurun_out/multi_c2_n4.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=4 value=%.0f step=%.3f e2e=%.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | grep "^{" | cut -c1-200
