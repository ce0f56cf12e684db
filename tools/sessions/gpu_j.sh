#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/raytracing-clj_b200/csrc/build/variants
for i in 1 2; do
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/j_noinline_$i.json 2>/dev/null
RTCLJ_LIB=$V/librtclj_ssinline.so timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/j_inline_$i.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/j_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],2), round(d['roofline']['frac'],4))
PY
