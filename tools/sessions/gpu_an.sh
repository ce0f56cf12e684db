#!/bin/bash
# A/B of build variants (VARIANTS, under build/variants): bench lines c3 (default kernel), lane kernel, c1, c2, c4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/an_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/an_smoke.log; exit 1; }
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/an_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/an_pytest.log
tail -n 3 gpurun_out/an_pytest.log
for v in ${VARIANTS:-new old}; do
  export RTCLJ_LIB=$PWD/raytracing-clj_b200/csrc/build/variants/librtclj_$v.so
  for w in ${WORKLOADS:-c3 lane c1 c2 c4}; do
    case $w in
      c3) args="--steps 4 --warmup 3";; lane) args="--kernel lane --steps 3 --warmup 3";; *) args="--workload $w --steps 5 --warmup 3";;
    esac
    timeout 300 python bench.py $args --no-cpu-baseline --no-extras > gpurun_out/an_${v}_$w.json 2> gpurun_out/an_${v}_$w.err
    python - $v $w <<'PY'
import json,sys
v,w=sys.argv[1:3]
try:
    d=json.loads(open("gpurun_out/an_%s_%s.json"%(v,w)).read().strip().splitlines()[-1])
    print(v, w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), d["workload_stats"]["rgb8_checksum"])
except Exception as e: print(v, w, "FAILED", e)
PY
  done
done
