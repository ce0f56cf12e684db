#!/bin/bash
# round-2 GPU session A: parity of the wavefront kernel + A/B timing + tuning variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 180 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1 || { echo 'SMOKE FAILED'; tail -5 gpurun_out/a_smoke.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python tools/quick_perf.py > gpurun_out/a_qp_wave.log 2>&1
RTCLJ_QP_LANE=1 timeout 300 python tools/quick_perf.py > gpurun_out/a_qp_lane.log 2>&1
RTCLJ_QP_STRICT=1 timeout 300 python tools/quick_perf.py cover_1920x1080x16 > gpurun_out/a_qp_strict.log 2>&1
for v in T640_S1280 T512_S1024 T512_S1536 T640_S768; do
  RTCLJ_LIB=$PWD/raytracing-clj_b200/csrc/build/variants/librtclj_$v.so timeout 300 python tools/quick_perf.py > gpurun_out/a_qp_$v.log 2>&1
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
