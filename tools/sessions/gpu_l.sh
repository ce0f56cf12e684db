#!/bin/bash
# final 1-GPU validation of the round: smoke, full GPU suite, the bench line, the reference arm, per-kernel lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/l_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/l_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
timeout 900 python bench.py > gpurun_out/l_bench_n1.json 2> gpurun_out/l_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/l_bench_reference.json 2> gpurun_out/l_bench_reference.err
for k in lane wave; do timeout 300 python bench.py --kernel $k --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/l_bench_$k.json 2> gpurun_out/l_bench_$k.err; done
python - <<'PY'
import json
for w in ["n1","reference","lane","wave"]:
    try:
        d=json.loads(open("gpurun_out/l_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],2), d.get("roofline",{}).get("frac"), d.get("e2e",{}).get("value"), (d.get("strict_order") or {}).get("value"))
    except Exception as e: print(w, "FAILED", e)
PY
