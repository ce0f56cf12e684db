#!/bin/bash
# quick check of a change to render_kernel: smoke, parity file, bench lines of the configs that run it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/aa_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/aa_smoke.log; exit 1; }
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/aa_pytest.log
tail -n 3 gpurun_out/aa_pytest.log
timeout 300 python bench.py --kernel lane --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/aa_bench_lane.json 2> gpurun_out/aa_bench_lane.err
for w in c1 c2 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/aa_bench_$w.json 2> gpurun_out/aa_bench_$w.err; done
timeout 300 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/aa_bench_c5.json 2> gpurun_out/aa_bench_c5.err
python - <<'PY'
import json
for w in ["lane","c1","c2","c4","c5"]:
    try:
        d=json.loads(open("gpurun_out/aa_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4))
    except Exception as e: print(w, "FAILED", e)
PY
