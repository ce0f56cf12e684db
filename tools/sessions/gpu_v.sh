#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
V=$PWD/raytracing-clj_b200/csrc/build/variants
for v in main olddiv nopref main olddiv nopref; do
  LIB=$PWD/raytracing-clj_b200/librtclj_b200.so; [ $v != main ] && LIB=$V/librtclj_$v.so
  RTCLJ_LIB=$LIB timeout 300 python bench.py --workload c5 --spp 16 --steps 4 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5/$v', round(d['ms_per_step'],2), round(d['roofline']['frac'],4))"
done
