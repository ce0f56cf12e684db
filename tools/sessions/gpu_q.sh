#!/bin/bash
# the bench line at 4 and at 2 GPUs (on a 4-GPU box), final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/q_bench_n$N.json 2> gpurun_out/q_bench_n$N.err
  tail -n 1 gpurun_out/q_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($N, round(d['value']/1e9,3), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,3), 'single', round(d['e2e_single_process']['value']/1e9,3), 'strict', round(d['strict_order']['value']/1e9,3), d['workload_stats']['rgb8_checksum'])"
done
