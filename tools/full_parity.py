"""Whole-image parity at the FULL size of BASELINE.json configs 2 and 3: the GPU render against the CPU
oracle over every pixel (not sampled rows).  The oracle needs minutes of host time for config 3
(2.76 G segments at ~12 M/s on 16 cores), so this is a tool, not a test.
Writes gpurun_out/r1_full_parity.json.  Usage: python tools/full_parity.py [2] [3]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import render

S, CAM = R.scenes, R.camera
which = [int(a) for a in sys.argv[1:]] or [2, 3]
cases = {2: ("2 (material scene 1080p, 100 spp)", S.main_hittables(), CAM.main_camera(1920), 100),
         3: ("3 (RTIOW cover scene 1080p, 500 spp)", S.cover_hittables(7), CAM.main_camera(1920, 1080, **S.COVER_CAMERA), 500)}
out = []
for c in which:
    name, world, cam, spp = cases[c]
    lin, rgb, st = render.render(world, cam, spp, 50, seed=1, flags=O.FLAGS_MAIN)
    t0 = time.perf_counter()
    lin_o, rgb_o, st_o = O.render(S.to_soa(world), cam, spp, 50, seed=1, flags=O.FLAGS_MAIN, threads=os.cpu_count(),
                                  samples_per_unit=st["samples_per_unit"])
    dt = time.perf_counter() - t0
    e = {"config": name, "image": f"{cam.width}x{cam.height}", "spp": spp, "n_spheres": len(world),
         "segments_gpu": st["segments"], "segments_oracle": int(st_o.segments),
         "linear_values_differing": int((lin != lin_o).sum()), "rgb8_values_differing": int((rgb != rgb_o).sum()),
         "values_compared": int(lin.size), "gpu_kernel_ms": round(st["kernel_ms"], 2), "oracle_seconds": round(dt, 1),
         "oracle_threads": os.cpu_count(), "samples_per_unit": st["samples_per_unit"]}
    print(json.dumps(e), flush=True)
    out.append(e)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r1_full_parity.json"), "w") as f:
    json.dump(out, f, indent=1)
