#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for spu in 0 14 54 100 167; do
  timeout 300 python bench.py --spu $spu --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/m_spu$spu.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/m_spu*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['workload_stats']['samples_per_unit'], round(d['ms_per_step'],2), round(d['roofline']['frac'],4))
PY
