"""Scratch probe: how evenly do the 8 shards of the bench frame split, and how long is the kernel's tail?
Renders every shard of an 8-GPU job in turn on ONE GPU (device-resident) and prints per-shard device times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render


def main():
    S, CAM = R.scenes, R.camera
    world = S.cover_hittables(7)
    cam = CAM.main_camera(1920, 1080, **S.COVER_CAMERA)
    ctx = render.Context(0)
    ctx.set_scene(world)
    out = torch.zeros((cam.height, cam.width, 3), dtype=torch.float64, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream

    def one(shard, spu=0, reps=2):
        best = None
        for _ in range(reps):
            ctx.render(cam, 500, 50, flags=_abi.FLAGS_MAIN, d_out_linear=out.data_ptr(), stream=stream, shard=shard, samples_per_unit=spu)
            st = ctx.stats(stream)
            if best is None or st["device_ms"] < best["device_ms"]:
                best = st
        return best

    one(None, reps=1)  # warm
    whole = one(None)
    print(json.dumps({"whole_ms": round(whole["device_ms"], 3), "segments": whole["segments"], "spu": whole["samples_per_unit"]}), flush=True)
    for count in (8,):
        for rows in (1, 2, 4, 8, 16):
            ms, seg = [], []
            for i in range(count):
                st = one((i, count, rows))
                ms.append(st["device_ms"]); seg.append(st["segments"])
            print(json.dumps({"count": count, "rows": rows, "max_ms": round(max(ms), 3), "mean_ms": round(sum(ms) / count, 3),
                              "ideal_ms": round(whole["device_ms"] / count, 3),
                              "seg_max_over_mean": round(max(seg) * count / sum(seg), 4),
                              "ms": [round(x, 2) for x in ms]}), flush=True)
    for spu in (10, 16, 20, 27, 36, 50):
        ms = [one((i, 8, 4), spu)["device_ms"] for i in (0, 3)]
        print(json.dumps({"rows": 4, "spu": spu, "ms": [round(x, 3) for x in ms]}), flush=True)
    for count in (2, 4, 16, 32):
        st = one((0, count, 4))
        print(json.dumps({"count": count, "shard0_ms": round(st["device_ms"], 3), "x_count": round(st["device_ms"] * count, 2),
                          "seg_share": round(st["segments"] * count / whole["segments"], 4)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
