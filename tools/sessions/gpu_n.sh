#!/bin/bash
# split kernel: a guarded first run (hang -> timeout kills the process), then parity tests and A/B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 60 python - > gpurun_out/n_first.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests")
import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render
world = R.scenes.cover_hittables(7)
cam = R.camera.main_camera(96, 54, **R.scenes.COVER_CAMERA)
lin, rgb, st = render.render(world, cam, 4, 50, seed=1, flags=_abi.FLAGS_MAIN | _abi.F_SPLIT_KERNEL, samples_per_unit=4)
lin_o, rgb_o, st_o = O.render(R.scenes.to_soa(world), cam, 4, 50, seed=1, flags=O.FLAGS_MAIN, threads=4)
print("first run:", np.array_equal(lin, lin_o), st["segments"], st_o.segments, st["device_ms"])
PY
echo "first rc=$?" >> gpurun_out/n_first.log
grep -q "first run: True" gpurun_out/n_first.log || { echo "SPLIT KERNEL FAILED ITS FIRST RUN"; cat gpurun_out/n_first.log | tail -n 5; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q -k "small_scene_kernel or two_scenes or strict_order_on_long" > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
RTCLJ_QP_KERNEL=split timeout 300 python tools/quick_perf.py > gpurun_out/n_qp_split.log 2>&1
for k in split lane2; do timeout 300 python bench.py --kernel $k --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/n_bench_$k.json 2> gpurun_out/n_bench_$k.err; done
python - <<'PY'
import json
for w in ["split","lane2"]:
    try:
        d=json.loads(open("gpurun_out/n_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"]/1e9,4), round(d["ms_per_step"],2), round(d["roofline"]["frac"],4))
    except Exception as e: print(w, "FAILED", e)
PY
