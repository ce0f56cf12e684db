#!/bin/bash
# round-2 GPU session C: wave kernel v3 (smaller code, compact scheduler): parity, A/B, variants, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -5 gpurun_out/c_smoke.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
timeout 300 python tools/quick_perf.py > gpurun_out/c_qp_wave.log 2>&1
for v in T640_S1280 T512_S1024 T768_S1024; do
  RTCLJ_LIB=$PWD/raytracing-clj_b200/csrc/build/variants/librtclj_$v.so timeout 300 python tools/quick_perf.py > gpurun_out/c_qp_$v.log 2>&1
done
for c in default_1920x1080x16 cover_1920x1080x16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave -s 2 -c 1 -f -o gpurun_out/c_wave_$c python tools/quick_perf.py $c 3 > gpurun_out/c_ncu_$c.log 2>&1
done
