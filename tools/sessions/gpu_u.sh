#!/bin/bash
# dedicated-cull-warp kernel with its uniform table loads intact: take-what-is-there vs wait-for-full-passes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/raytracing-clj_b200/csrc/build/variants
timeout 600 python -m pytest tests -m gpu -x -q -k "small_scene_kernel or two_scenes or strict_order_on_long" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log
for v in main splitwf; do
  LIB=$PWD/raytracing-clj_b200/librtclj_b200.so; [ $v = splitwf ] && LIB=$V/librtclj_splitwf.so
  RTCLJ_LIB=$LIB timeout 300 python bench.py --kernel split --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('split/$v', round(d['value']/1e9,4), round(d['ms_per_step'],2), round(d['roofline']['frac'],4))"
  RTCLJ_LIB=$LIB RTCLJ_QP_KERNEL=split timeout 200 python tools/quick_perf.py 2>/dev/null | grep -o '"case": "[^"]*"\|"ms": [0-9.]*' | paste - - | tr '\n' ' '; echo
done
