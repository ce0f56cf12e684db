#!/bin/bash
# hygiene: every kernel and mode on tiny renders against the oracle, and a fuzz soak that rotates the kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render
S, CAM = R.scenes, R.camera
world, cam = S.cover_hittables(7), CAM.main_camera(48, 27, **S.COVER_CAMERA)
lin_o, rgb_o, st_o = O.render(S.to_soa(world), cam, 3, 50, seed=1, flags=O.FLAGS_MAIN, threads=4, samples_per_unit=3)
for name, extra in (("lane", _abi.F_LANE_KERNEL), ("lane2", _abi.F_LANE2_KERNEL), ("wave", _abi.F_WAVE_KERNEL),
                    ("split", _abi.F_SPLIT_KERNEL), ("smem-table", _abi.F_SMEM_TABLE)):
    lin, rgb, st = render.render(world, cam, 3, 50, seed=1, flags=_abi.FLAGS_MAIN | extra, samples_per_unit=3)
    print(name, bool(np.array_equal(lin, lin_o)), st["segments"] == st_o.segments, flush=True)
# strict order through the sample buffer, chunked units with the wavefront kernel's last-arriver finish, render to PPM text
lin_s, _, _ = O.render(S.to_soa(world), cam, 64, 50, seed=2, flags=O.FLAGS_MAIN, threads=4, samples_per_unit=64)
lin, _, _ = render.render(world, cam, 64, 50, seed=2, flags=_abi.FLAGS_MAIN | _abi.F_LANE2_KERNEL, samples_per_unit=64)
print("strict buffer", bool(np.array_equal(lin, lin_s)), flush=True)
lin_c, _, _ = O.render(S.to_soa(world), cam, 9, 50, seed=3, flags=O.FLAGS_MAIN, threads=4, samples_per_unit=2)
lin, rgb, _ = render.render(world, cam, 9, 50, seed=3, flags=_abi.FLAGS_MAIN | _abi.F_WAVE_KERNEL, samples_per_unit=2)
print("wave chunked", bool(np.array_equal(lin, lin_c)), flush=True)
text, _ = render.render_ppm(world, cam, 9, 50, seed=3, flags=_abi.FLAGS_MAIN, samples_per_unit=2)
print("render_ppm", text == render.encode_ppm(rgb), flush=True)
PY
# (compute-sanitizer is closed on this pool: the script runs plain, as a functional check of every kernel and mode)
timeout 300 python /tmp/san.py > gpurun_out/p_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/p_kernels.log
RTCLJ_FUZZ_CASES=${RTCLJ_SOAK:-6000} timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fuzz > gpurun_out/p_fuzz.log 2>&1; echo "fuzz rc=$?" >> gpurun_out/p_fuzz.log
