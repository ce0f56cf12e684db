#!/bin/bash
# round-2 GPU session I: strict order through the sample buffer: tests + bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/i_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/i_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err
for w in c1 c2; do timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/i_bench_$w.json 2> gpurun_out/i_bench_$w.err; done
