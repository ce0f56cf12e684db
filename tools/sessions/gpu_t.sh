#!/bin/bash
# final record of the round: full GPU suite, fuzz soak, bench lines for every workload, reference arm, all-configs report
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/t_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log
RTCLJ_FUZZ_CASES=3000 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > gpurun_out/t_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/t_fuzz.log
timeout 900 python bench.py > gpurun_out/t_bench_n1.json 2> gpurun_out/t_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t_bench_reference.json 2> gpurun_out/t_bench_reference.err
for k in lane wave split; do timeout 300 python bench.py --kernel $k --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/t_bench_$k.json 2> gpurun_out/t_bench_$k.err; done
for w in c1 c2 c4; do timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/t_bench_$w.json 2> gpurun_out/t_bench_$w.err; done
timeout 900 python bench.py --workload c5 --steps 2 --warmup 3 --no-extras > gpurun_out/t_bench_c5.json 2> gpurun_out/t_bench_c5.err
timeout 1500 python tools/report_configs.py > gpurun_out/t_configs.log 2>&1
python - <<'PY'
import json
for w in ["n1","reference","lane","wave","split","c1","c2","c4","c5"]:
    try:
        d=json.loads(open("gpurun_out/t_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), (d.get("roofline") or {}).get("frac"), "e2e", round(d["e2e"]["value"]/1e9,4), "strict", (d.get("strict_order") or {}).get("value"))
    except Exception as e: print(w, "FAILED", e)
PY
