#!/bin/bash
# A/B of two builds on the short renders of tools/quick_perf.py (the library's own kernel choice)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in ${VARIANTS:-new old new old}; do
  export RTCLJ_LIB=$PWD/raytracing-clj_b200/csrc/build/variants/librtclj_$v.so
  echo "== $v"
  for c in cover_480x270x16 cover_1920x1080x16 cover_normalshade_1920x1080x32 default_1920x1080x16 realm_1920x1080x16; do
    timeout 120 python tools/quick_perf.py $c 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['case'], d['ms'])"
  done
done > gpurun_out/al_quick.log 2>&1
cat gpurun_out/al_quick.log
