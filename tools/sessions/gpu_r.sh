#!/bin/bash
# after a change to shared arithmetic helpers: full GPU suite, fuzz soak, bench lines (default, lane, c1, c2, c4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/r_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/r_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
RTCLJ_FUZZ_CASES=${RTCLJ_SOAK:-3000} timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > gpurun_out/r_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/r_fuzz.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_n1.json 2> gpurun_out/r_bench_n1.err
timeout 300 python bench.py --kernel lane --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r_bench_lane.json 2> gpurun_out/r_bench_lane.err
for w in c1 c2 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r_bench_$w.json 2> gpurun_out/r_bench_$w.err; done
python - <<'PY'
import json
for w in ["n1","lane","c1","c2","c4"]:
    try:
        d=json.loads(open("gpurun_out/r_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4))
    except Exception as e: print(w, "FAILED", e)
PY
