#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_host_paths.py -m gpu -x -q > gpurun_out/k_pytest_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest_$N.log
timeout 600 python tools/bench_render_ppm.py $N > gpurun_out/k_render_ppm_n$N.log 2>&1
timeout 600 python tools/bench_render_ppm.py 1 > gpurun_out/k_render_ppm_n1.log 2>&1
