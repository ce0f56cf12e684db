#!/bin/bash
# exhaustive fp64 scan vs fp32 cull on scenes of 1..12 spheres; bench lines of the small configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/nocull_crossover.py > gpurun_out/y_crossover.log 2>&1; cat gpurun_out/y_crossover.log
timeout 300 python bench.py --kernel lane --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/y_bench_lane.json 2> gpurun_out/y_bench_lane.err
for w in c1 c2 c4; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/y_bench_$w.json 2> gpurun_out/y_bench_$w.err; done
python - <<'PY'
import json
for w in ["lane","c1","c2","c4"]:
    try:
        d=json.loads(open("gpurun_out/y_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4))
    except Exception as e: print(w, "FAILED", e)
PY
