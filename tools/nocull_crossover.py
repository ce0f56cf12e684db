"""Scratch probe: from how many spheres on does the fp32 cull pay?  Prefixes of the default scene, 1920x1080 x 16 spp,
with and without RTCLJ_F_NO_CULL (exhaustive fp64 scan), best of 3 device times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render

S, CAM = R.scenes, R.camera
full = S.main_hittables() + S.cover_hittables(7)[4:12]
cam = CAM.main_camera(1920)
out = torch.zeros((cam.height, cam.width, 3), dtype=torch.float64, device="cuda:0")
stream = torch.cuda.current_stream().cuda_stream
for n in (1, 2, 3, 4, 5, 6, 8, 12):
    world = full[:n]
    row = {"n": n}
    for name, extra in (("cull", 0), ("scan", _abi.F_NO_CULL)):
        for shade, fl in (("path", _abi.FLAGS_MAIN), ("normal", _abi.FLAGS_I)):
            ctx = render.Context(0)
            ctx.set_scene(world)
            best = 1e9
            for _ in range(3):
                ctx.render(cam, 16, 50, flags=fl | extra | _abi.F_LANE_KERNEL, d_out_linear=out.data_ptr(), stream=stream)
                best = min(best, ctx.stats(stream)["device_ms"])
            ctx.close()
            row[f"{shade}_{name}_ms"] = round(best, 3)
    print(json.dumps(row), flush=True)
