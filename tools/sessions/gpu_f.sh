#!/bin/bash
# round-2 GPU session F: full GPU test-suite with the new host-path tests, the default bench line, a lane2 variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nproc > gpurun_out/f_nproc.txt
timeout 180 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/f_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
V=$PWD/raytracing-clj_b200/csrc/build/variants
RTCLJ_LIB=$V/librtclj_l2sumsg.so timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --kernel lane2 > gpurun_out/f_bench_l2sumsg.json 2> gpurun_out/f_bench_l2sumsg.err
timeout 900 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
timeout 300 python tools/bench_gather.py 1 > gpurun_out/f_gather1.log 2>&1
