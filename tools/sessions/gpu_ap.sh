#!/bin/bash
# break-even of one vs two paths per lane on the cover scene at 1920x1080: 16 / 24 / 32 / 48 spp (3.3e7 ... 1e8 samples)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for spp in 16 24 32 48; do
  for k in lane lane2; do
    RTCLJ_QP_SPP=$spp RTCLJ_QP_KERNEL=$k timeout 120 python tools/quick_perf.py cover_1920x1080x16 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($spp, '$k', d['ms'])"
  done
done | tee gpurun_out/ap_breakeven.log
