#!/bin/bash
# primary-ray kernel: tuning variants (tools/build_variants.sh) on config 4 and on a 5-sphere depth-1 render
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in ${VARIANTS:-base mb3 mb5 unroll unroll3 t128}; do
  export RTCLJ_LIB=$PWD/raytracing-clj_b200/csrc/build/variants/librtclj_$v.so
  timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_c4_$v.json 2> gpurun_out/ab_c4_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/ab_c4_%s.json"%v).read().strip().splitlines()[-1])
    print(v, round(d["value"]/1e9,3), round(d["ms_per_step"],3), d["workload_stats"]["rgb8_checksum"])
except Exception as e: print(v, "FAILED", e)
PY
done
