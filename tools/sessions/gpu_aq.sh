#!/bin/bash
# last check of the round: smoke, full GPU suite, the default bench line (short)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/aq_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/aq_smoke.log; exit 1; }
tail -n 2 gpurun_out/aq_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/aq_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/aq_pytest.log
tail -n 3 gpurun_out/aq_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/aq_bench_n1.json 2> gpurun_out/aq_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/aq_bench_n1.json').read().strip().splitlines()[-1])
print('n1', round(d['value']/1e9,4), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['gpu_launches'])"
RTCLJ_QP_SPP=32 timeout 100 python tools/quick_perf.py cover_1920x1080x16 3 | tail -1 | cut -c1-120
