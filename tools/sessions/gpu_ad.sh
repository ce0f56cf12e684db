#!/bin/bash
# ticket order experiment: shard probe + bench lines (no test suite)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/ad_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/ad_smoke.log; exit 1; }
timeout 600 python tools/shard_balance.py > gpurun_out/ad_shards.log 2>&1; cat gpurun_out/ad_shards.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ad_bench_n1.json 2> gpurun_out/ad_bench_n1.err
timeout 300 python bench.py --kernel lane --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ad_bench_lane.json 2> gpurun_out/ad_bench_lane.err
for w in c1 c2; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ad_bench_$w.json 2> gpurun_out/ad_bench_$w.err; done
python - <<'PY'
import json
for w in ["n1","lane","c1","c2"]:
    try:
        d=json.loads(open("gpurun_out/ad_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["workload_stats"]["samples_per_unit"])
    except Exception as e: print(w, "FAILED", e)
PY
