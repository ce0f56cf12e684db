#!/bin/bash
# per-warp timestamps of the two-paths kernel (tail probe build)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export RTCLJ_LIB=$PWD/raytracing-clj_b200/csrc/build/variants/librtclj_probe.so
for c in 8 1 32; do timeout 200 python tools/tail_probe.py $c 4; done > gpurun_out/ae_tail.log 2>&1
cat gpurun_out/ae_tail.log
