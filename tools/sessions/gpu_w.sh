#!/bin/bash
# shard balance / tail probe on one GPU (tools/shard_balance.py)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/shard_balance.py > gpurun_out/w_shards.log 2>&1; echo "rc=$?" >> gpurun_out/w_shards.log
cat gpurun_out/w_shards.log
