#!/bin/bash
# round-2 GPU session E: A/B of the lane kernels on the FULL bench workload (and the quick cases)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/raytracing-clj_b200/csrc/build/variants
run() { # name, lib, kernel
  RTCLJ_LIB=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --kernel $3 > gpurun_out/e_bench_$1.json 2> gpurun_out/e_bench_$1.err
  RTCLJ_LIB=$2 RTCLJ_QP_KERNEL=$3 timeout 200 python tools/quick_perf.py > gpurun_out/e_qp_$1.log 2>&1
}
run lane_cands3 $PWD/raytracing-clj_b200/librtclj_b200.so lane
run lane_cands2 $V/librtclj_cands2.so lane
run lane2 $PWD/raytracing-clj_b200/librtclj_b200.so lane2
run lane2_t512 $V/librtclj_l2t512.so lane2
run lane2_t576 $V/librtclj_l2t576.so lane2
run lane2_philox $V/librtclj_l2philox.so lane2
for f in gpurun_out/e_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('e_bench_')[1], round(d['value']/1e9,4), 'G rays/s', round(d['ms_per_step'],2), 'ms', 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e9,4), d['e2e']['host_buffers'])
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
