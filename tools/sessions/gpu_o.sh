#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
for k in lane lane2 lane lane2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --kernel $k --no-extras 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$k', round(d['value']/1e9,3), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value']/1e9,3))"
done
