#!/bin/bash
# validation of a build: smoke, full GPU suite, fuzz soak, bench lines (default, lane kernel, c1, c2, c4), short renders
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
[ -x ./gpurun_div_check ] && { timeout 200 ./gpurun_div_check | tee gpurun_out/am_div_check.log; }
timeout 180 python __graft_entry__.py smoke > gpurun_out/am_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/am_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/am_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/am_pytest.log
tail -n 3 gpurun_out/am_pytest.log
RTCLJ_FUZZ_CASES=${RTCLJ_SOAK:-3000} timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > gpurun_out/am_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/am_fuzz.log
tail -n 2 gpurun_out/am_fuzz.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/am_bench_n1.json 2> gpurun_out/am_bench_n1.err
timeout 300 python bench.py --kernel lane --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/am_bench_lane.json 2> gpurun_out/am_bench_lane.err
for w in c1 c2 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/am_bench_$w.json 2> gpurun_out/am_bench_$w.err; done
python - <<'PY'
import json
for w in ["n1","lane","c1","c2","c4"]:
    try:
        d=json.loads(open("gpurun_out/am_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["e2e"]["value"] if d.get("e2e") else None, (d.get("strict_order") or {}).get("ms_per_step"))
    except Exception as e: print(w, "FAILED", e)
PY
for c in cover_480x270x16 cover_1920x1080x16 cover_normalshade_1920x1080x32 default_1920x1080x16 realm_1920x1080x16 i_3840x2160x16 field10k_960x540x4; do
  timeout 120 python tools/quick_perf.py $c 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['case'], d['ms'], d['exact_per_seg'], d['pref_per_seg'])"
done
