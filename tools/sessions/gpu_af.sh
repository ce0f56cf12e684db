#!/bin/bash
# bottom-rows-first ticket order: full GPU suite, fuzz soak, bench lines incl. extras
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/af_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/af_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/af_pytest.log
tail -n 3 gpurun_out/af_pytest.log
RTCLJ_FUZZ_CASES=${RTCLJ_SOAK:-1500} timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > gpurun_out/af_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/af_fuzz.log
tail -n 2 gpurun_out/af_fuzz.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/af_bench_n1.json 2> gpurun_out/af_bench_n1.err
for w in c1 c2 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/af_bench_$w.json 2> gpurun_out/af_bench_$w.err; done
python - <<'PY'
import json
for w in ["n1","c1","c2","c4"]:
    try:
        d=json.loads(open("gpurun_out/af_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["e2e"]["value"], (d.get("strict_order") or {}).get("ms_per_step"))
    except Exception as e: print(w, "FAILED", e)
PY
