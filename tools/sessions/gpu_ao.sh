#!/bin/bash
# the record of the round's final build: validation (smoke, GPU suite, fuzz soak), every bench line, ncu of the bench
# kernel and of the primary-ray kernel, launch lists, all five configs, short renders
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/ao_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/ao_smoke.log; exit 1; }
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/ao_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ao_pytest.log
tail -n 3 gpurun_out/ao_pytest.log
RTCLJ_FUZZ_CASES=${RTCLJ_SOAK:-3000} timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > gpurun_out/ao_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/ao_fuzz.log
tail -n 2 gpurun_out/ao_fuzz.log
timeout 900 python bench.py > gpurun_out/ao_bench_n1.json 2> gpurun_out/ao_bench_n1.err
timeout 300 python bench.py --kernel lane --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ao_bench_lane.json 2> gpurun_out/ao_bench_lane.err
for w in c1 c2 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/ao_bench_$w.json 2> gpurun_out/ao_bench_$w.err; done
timeout 400 python bench.py --workload c5 --steps 2 --warmup 1 --no-extras > gpurun_out/ao_bench_c5.json 2> gpurun_out/ao_bench_c5.err
python - <<'PY'
import json
for w in ["n1","lane","c1","c2","c4","c5"]:
    try:
        d=json.loads(open("gpurun_out/ao_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["e2e"]["value"] if d.get("e2e") else None, (d.get("strict_order") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(w, "FAILED", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ao_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ao_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_lane2 -s 3 -c 1 -f -o gpurun_out/ao_lane2_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ao_ncu_lane2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_primary -s 3 -c 1 -f -o gpurun_out/ao_primary_c4 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ao_ncu_c4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/ao_launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ao_ncu_launch_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/ao_lane_c2 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ao_ncu_c2.log 2>&1
timeout 1500 python tools/report_configs.py > gpurun_out/ao_configs.log 2>&1
tail -n 3 gpurun_out/ao_configs.log
for c in cover_480x270x16 cover_1920x1080x16 cover_normalshade_1920x1080x32 default_1920x1080x16 realm_1920x1080x16 i_3840x2160x16 field10k_960x540x4; do
  timeout 120 python tools/quick_perf.py $c 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['case'], d['ms'], d['exact_per_seg'], d['pref_per_seg'])"
done | tee gpurun_out/ao_quick.log
