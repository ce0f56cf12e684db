#!/bin/bash
# round-2 GPU session D: the three small-scene kernels: parity (default = lane2) and A/B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/d_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -5 gpurun_out/d_smoke.log; exit 1; }
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
for k in lane lane2 wave; do
  RTCLJ_QP_KERNEL=$k timeout 300 python tools/quick_perf.py > gpurun_out/d_qp_$k.log 2>&1
done
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_lane2 -s 2 -c 1 -f -o gpurun_out/d_lane2_cover python tools/quick_perf.py cover_1920x1080x16 3 > gpurun_out/d_ncu.log 2>&1
