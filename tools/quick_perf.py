"""Scratch perf probe (not the bench): device-resident renders of the cover scene."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render


def peaks():
    a, b, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int32()
    _abi.check(_abi.lib().rtclj_calibrate_peaks(0, C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
    return {"ffma_tflops": a.value, "ffma2_tflops": b.value, "dfma_tflops": c.value, "sms": n.value}


def run(name, world, cam, spp, depth, flags, reps=3, **kw):
    which = os.environ.get("RTCLJ_QP_KERNEL")
    if which:
        flags |= {"lane": _abi.F_LANE_KERNEL, "lane2": _abi.F_LANE2_KERNEL, "wave": _abi.F_WAVE_KERNEL, "split": _abi.F_SPLIT_KERNEL}[which]
        name += f"[{which}]"
    if os.environ.get("RTCLJ_QP_FLAGS"):
        flags |= int(os.environ["RTCLJ_QP_FLAGS"], 0)
        name += f"[+{os.environ['RTCLJ_QP_FLAGS']}]"
    if os.environ.get("RTCLJ_QP_STRICT"):
        kw["samples_per_unit"] = spp
        name += "[strict]"
    ctx = render.Context(0)
    ctx.set_scene(world)
    out = torch.zeros((cam.height, cam.width, 3), dtype=torch.float64, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps):
        ctx.render(cam, spp, depth, flags=flags, d_out_linear=out.data_ptr(), stream=stream, **kw)
        st = ctx.stats(stream)
        if best is None or st["device_ms"] < best["device_ms"]:
            best = st
    n = len(world)
    segs = best["segments"] / (best["device_ms"] * 1e-3)
    tf = segs * (17 * n + 5) / 1e12
    print(json.dumps({"case": name, "n": n, "ms": round(best["device_ms"], 3), "Gseg_s": round(segs / 1e9, 4),
                      "alg_TFLOPs": round(tf, 2), "seg_per_sample": round(best["segments"] / best["samples"], 3),
                      "exact_per_seg": round(best["exact_tests"] / max(1, best["segments"]), 2), "pref_per_seg": round(best["prefilter_tests"] / max(1, best["segments"]), 2),
                      "overflows": best["list_overflows"], "spu": best["samples_per_unit"]}), flush=True)
    ctx.close()


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else None
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    S, CAM = R.scenes, R.camera
    cover = S.cover_hittables(7)
    cases = {
        "cover_480x270x16": lambda: (cover, CAM.main_camera(480, 270, **S.COVER_CAMERA), 16),
        "cover_1920x1080x16": lambda: (cover, CAM.main_camera(1920, 1080, **S.COVER_CAMERA), 16),
        "cover_normalshade_1920x1080x32": lambda: (cover, CAM.main_camera(1920, 1080, **S.COVER_CAMERA), 32, _abi.FLAGS_I),
        "default_1920x1080x16": lambda: (S.main_hittables(), CAM.main_camera(1920), 16),
        "realm_1920x1080x16": lambda: (S.realm_hittables(), CAM.realm_camera(1920), 16, _abi.FLAGS_REALM),
        "i_3840x2160x16": lambda: (S.i_hittables(), CAM.i_camera(3840), 16, _abi.FLAGS_I),
        "field10k_960x540x4": lambda: (S.field_hittables(7), CAM.main_camera(960, 540, **S.FIELD_CAMERA), 4),
    }
    if not only:
        print(json.dumps(peaks()), flush=True)
    for name, mk in cases.items():
        if only and name != only:
            continue
        world, cam, spp, *fl = mk()
        spp = int(os.environ.get("RTCLJ_QP_SPP", spp))
        run(name, world, cam, spp, 50, fl[0] if fl else _abi.FLAGS_MAIN, reps=reps)
