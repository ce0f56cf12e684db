"""Scratch probe (needs a -DRTCLJ_TAIL_PROBE build, tools/build_variants.sh): per-warp timestamps of the two-paths
kernel -- start, first time a lane found the queue empty, end -- on one shard of the bench frame.
usage: RTCLJ_LIB=.../librtclj_probe.so python tools/tail_probe.py [shard_count=8] [shard_rows=4]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render

count = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S, CAM = R.scenes, R.camera
world = S.cover_hittables(7)
cam = CAM.main_camera(1920, 1080, **S.COVER_CAMERA)
ctx = render.Context(0)
ctx.set_scene(world)
out = torch.zeros((cam.height, cam.width, 3), dtype=torch.float64, device="cuda:0")
stream = torch.cuda.current_stream().cuda_stream
path = "/tmp/rtclj_tail_probe.bin"
os.environ["RTCLJ_TAIL_PROBE_FILE"] = path
shard = (0, count, rows) if count > 1 else None
for rep in range(2):
    ctx.render(cam, 500, 50, flags=_abi.FLAGS_MAIN | _abi.F_LANE2_KERNEL, d_out_linear=out.data_ptr(), stream=stream, shard=shard)
    st = ctx.stats(stream)
rec = np.fromfile(path, dtype=np.uint64).reshape(-1, 4)
t0 = rec[:, 0].min()
start, dry, end = (rec[:, 0] - t0) / 1e6, (rec[:, 1] - t0) / 1e6, (rec[:, 2] - t0) / 1e6
it, it_dry = (rec[:, 3] >> np.uint64(32)).astype(np.int64), (rec[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
q = lambda a: [round(float(x), 3) for x in np.percentile(a, [0, 5, 50, 95, 100])]
print(json.dumps({"shards": count, "device_ms": round(st["device_ms"], 3), "warps": len(rec),
                  "start_ms_pctl": q(start), "first_dry_ms_pctl": q(dry), "end_ms_pctl": q(end),
                  "dry_to_end_ms_pctl": q(end - dry), "queue_empty_at_ms": round(float(dry.min()), 3),
                  "last_end_ms": round(float(end.max()), 3),
                  "idle_warp_ms_after_queue_empty": round(float((end.max() - end).mean()), 3),
                  "iterations_pctl": q(it), "iterations_after_dry_pctl": q(it - it_dry),
                  "us_per_iteration_before_dry": round(float(((dry - start) * 1e3 / np.maximum(it_dry, 1)).mean()), 2),
                  "us_per_iteration_after_dry": round(float(((end - dry) * 1e3 / np.maximum(it - it_dry, 1)).mean()), 2)}))
ctx.close()
