#!/bin/bash
# ncu (source-level) of render_kernel<true> on config 2 (5 spheres, 1920x1080 x 100 spp)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/aj_plain_c2.json 2> gpurun_out/aj_plain_c2.err || { echo 'bench failed'; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/aj_lane_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/aj_ncu_c2.log 2>&1
ls -la gpurun_out/aj_*
