#!/bin/bash
# A/B of two builds (librtclj_new.so / librtclj_old.so under build/variants): parity + fuzz on the default build, then bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/ak_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -n 5 gpurun_out/ak_smoke.log; exit 1; }
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ak_pytest.log
tail -n 3 gpurun_out/ak_pytest.log
RTCLJ_FUZZ_CASES=${RTCLJ_SOAK:-2000} timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > gpurun_out/ak_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/ak_fuzz.log
tail -n 2 gpurun_out/ak_fuzz.log
for v in new old; do
  export RTCLJ_LIB=$PWD/raytracing-clj_b200/csrc/build/variants/librtclj_$v.so
  for w in c3 lane c1 c2; do
    case $w in
      c3) args="--steps 4 --warmup 3";; lane) args="--kernel lane --steps 3 --warmup 3";; *) args="--workload $w --steps 5 --warmup 3";;
    esac
    timeout 300 python bench.py $args --no-cpu-baseline --no-extras > gpurun_out/ak_${v}_$w.json 2> gpurun_out/ak_${v}_$w.err
    python - $v $w <<'PY'
import json,sys
v,w=sys.argv[1:3]
try:
    d=json.loads(open("gpurun_out/ak_%s_%s.json"%(v,w)).read().strip().splitlines()[-1])
    print(v, w, round(d["value"]/1e9,4), round(d["ms_per_step"],3), d["workload_stats"]["rgb8_checksum"])
except Exception as e: print(v, w, "FAILED", e)
PY
  done
done
