"""End-to-end "render to the text of scene.ppm" (src/raytracing.clj:141-175: render loop + write-color! loop)
on N GPUs, one process: (a) rtclj_render_multi into host memory + the host P3 writer, (b) the same + the device
P3 writer through host buffers (rtclj_encode_ppm_p3_gpu), (c) rtclj_render_multi_ppm (shards assembled on GPU 0
by peer copies, device P3 writer there, only the text crosses PCIe).  Wall clock, best of 5.
usage: python tools/bench_render_ppm.py [n_gpus]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render

S, CAM = R.scenes, R.camera


def best(fn, reps=5):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        t.append(time.perf_counter() - t0)
    return min(t) * 1e3, r


def main():
    n = C.c_int()
    _abi.check(_abi.lib().rtclj_device_count(C.byref(n)))
    ngpu = min(int(sys.argv[1]) if len(sys.argv) > 1 else n.value, n.value)
    devs = list(range(ngpu))
    rows = []
    for label, world, cam, spp, depth, flags in (
            ("raytracing-i 3840x2160, 100 spp", S.i_hittables(), CAM.i_camera(3840), 100, 50, _abi.FLAGS_I),
            ("main scene 1920x1080, 100 spp", S.main_hittables(), CAM.main_camera(1920), 100, 50, _abi.FLAGS_MAIN)):
        soa = S.to_soa(world)
        lib = _abi.lib()
        H, W = cam.height, cam.width
        sc, cm = render._scene_struct(soa), render._camera_struct(cam)
        prm = _abi.Params(spp, depth, 1, flags, 0, 0, 0, 0, 0, 0)
        arr = (C.c_int32 * ngpu)(*devs)
        cap = 64 + W * H * 12
        p_rgb, p_txt = C.c_void_p(), C.c_void_p()   # caller buffers allocated once, pinned: the C ABI as a host would use it
        _abi.check(lib.rtclj_host_alloc(W * H * 3, C.byref(p_rgb)))
        _abi.check(lib.rtclj_host_alloc(cap, C.byref(p_txt)))
        ln, st = C.c_size_t(), _abi.Stats()

        def host_path(device_writer):
            _abi.check(lib.rtclj_render_multi(C.byref(sc), C.byref(cm), C.byref(prm), arr, ngpu, None, p_rgb, C.byref(st)))
            if device_writer:
                _abi.check(lib.rtclj_encode_ppm_p3_gpu(0, p_rgb, W, H, p_txt, cap, C.byref(ln)))
            else:
                _abi.check(lib.rtclj_encode_ppm_p3(p_rgb, W, H, p_txt, cap, C.byref(ln)))
            return C.string_at(p_txt, ln.value) if check[0] else ln.value

        def fused():
            _abi.check(lib.rtclj_render_multi_ppm(C.byref(sc), C.byref(cm), C.byref(prm), arr, ngpu, p_txt, cap, C.byref(ln), C.byref(st)))
            return C.string_at(p_txt, ln.value) if check[0] else ln.value

        check = [True]
        a, b, c = host_path(False), host_path(True), fused()
        assert a == b == c
        check[0] = False
        a_ms, _ = best(lambda: host_path(False))
        b_ms, _ = best(lambda: host_path(True))
        c_ms, _ = best(fused)
        st = st.as_dict()
        lib.rtclj_host_free(p_rgb); lib.rtclj_host_free(p_txt)
        rows.append({"scope": label, "n_gpus": ngpu, "text_MB": round(len(a) / 1e6, 1), "kernel_ms": round(st["kernel_ms"], 2),
                     "render_multi+host_p3_ms": round(a_ms, 2), "render_multi+device_p3_via_host_ms": round(b_ms, 2),
                     "render_multi_ppm_ms": round(c_ms, 2)})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
