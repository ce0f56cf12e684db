#!/bin/bash
# tiny scenes with and without the fp32 cull (RTCLJ_F_NO_CULL = 0x10000)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for c in default_1920x1080x16 realm_1920x1080x16 i_3840x2160x16; do
  timeout 120 python tools/quick_perf.py $c 3
  RTCLJ_QP_FLAGS=0x10000 timeout 120 python tools/quick_perf.py $c 3
done > gpurun_out/x_nocull.log 2>&1
cat gpurun_out/x_nocull.log
