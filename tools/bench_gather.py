"""Gather of the shards of an image rendered on N GPUs (north_star: "gathered to the host, with NCCL used
only if a device-side gather beats the host copy"; the reference's gather is the concat of its two futures,
src/raytracing.clj:168-171).  Rows are interleaved over the GPUs (1-row tiles, as rtclj_render_multi and
bench.py do), every GPU holds a full-size image with its own rows filled.  Measured, wall clock, one process:

  host      every GPU copies its rows straight into ONE pinned host framebuffer (N strided 2-D copies in
            parallel, one per GPU) -- what the library does
  peer      every GPU copies its rows into GPU 0's image over NVLink (N-1 strided peer copies), then GPU 0
            copies the assembled image to the host in one piece
  peer+p3   as `peer`, then the device P3 writer (rtclj_ctx_encode_ppm_p3) runs on the assembled 8-bit image
            on GPU 0 and only the text goes to the host -- the one thing a device-side gather makes possible
  host+p3   the host gather of the 8-bit image followed by the host P3 writer (rtclj_encode_ppm_p3)

usage: python tools/bench_gather.py [n_gpus]   -> JSON lines + profiles/r2_gather.md when run under gpurun"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cuda.bindings import runtime as rt

from raytracing_clj_b200 import _abi, render


def ck(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA error {err}")
    return res[1:] if len(res) > 1 else None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    n = min(n, torch.cuda.device_count())
    lib = _abi.lib()
    for a in range(n):
        torch.cuda.set_device(a)
        for b in range(n):
            if a != b:
                rt.cudaDeviceEnablePeerAccess(b, 0)  # already-enabled is fine
    rows = []
    for name, (W, H) in {"1920x1080": (1920, 1080), "3840x2160": (3840, 2160)}.items():
        for what, elem in (("linear f64", 8), ("rgb8", 1)):
            row_bytes = W * 3 * elem
            img_bytes = H * row_bytes
            dev = [torch.full((img_bytes,), 1 + d, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
            if what == "rgb8":  # a plausible value distribution for the P3 writer
                g = torch.Generator(device="cuda:0").manual_seed(1)
                base = (torch.rand(img_bytes, device="cuda:0", generator=g) ** 0.5 * 255).to(torch.uint8)
                for d in range(n):
                    dev[d].copy_(base.to(f"cuda:{d}"))
            streams = [torch.cuda.Stream(device=f"cuda:{d}") for d in range(n)]
            hp = C.c_void_p()
            _abi.check(lib.rtclj_host_alloc(img_bytes, C.byref(hp)))
            text_cap = 64 + W * H * 12
            tp = C.c_void_p()
            _abi.check(lib.rtclj_host_alloc(text_cap, C.byref(tp)))
            d_text = torch.empty(text_cap, dtype=torch.uint8, device="cuda:0") if what == "rgb8" else None
            ctx0 = render.Context(0)

            def sync():
                for d in range(n):
                    torch.cuda.synchronize(d)

            def rows_copy(dst_ptr, d, kind):  # GPU d's rows d, d+n, d+2n ... as one strided 2-D copy
                cnt = len(range(d, H, n))
                off = d * row_bytes
                ck(rt.cudaMemcpy2DAsync(dst_ptr + off, n * row_bytes, dev[d].data_ptr() + off, n * row_bytes,
                                        row_bytes, cnt, kind, streams[d].cuda_stream))

            def host_gather():
                for d in range(n):
                    torch.cuda.set_device(d)
                    rows_copy(hp.value, d, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost)
                sync()

            def peer_gather(to_host=True):
                for d in range(1, n):
                    torch.cuda.set_device(d)
                    rows_copy(dev[0].data_ptr(), d, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
                sync()
                torch.cuda.set_device(0)
                if to_host:
                    ck(rt.cudaMemcpyAsync(hp.value, dev[0].data_ptr(), img_bytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost,
                                          streams[0].cuda_stream))
                    torch.cuda.synchronize(0)

            def peer_p3():
                peer_gather(to_host=False)
                torch.cuda.set_device(0)
                nbytes = ctx0.encode_ppm(dev[0].data_ptr(), W, H, d_text.data_ptr(), text_cap, streams[0].cuda_stream)
                ck(rt.cudaMemcpyAsync(tp.value, d_text.data_ptr(), nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost,
                                      streams[0].cuda_stream))
                torch.cuda.synchronize(0)
                return nbytes

            def host_p3():
                host_gather()
                ln = C.c_size_t()
                _abi.check(lib.rtclj_encode_ppm_p3(hp, W, H, tp, text_cap, C.byref(ln)))
                return ln.value

            def best(fn, reps=5):
                fn()
                t = []
                for _ in range(reps):
                    sync()
                    t0 = time.perf_counter()
                    fn()
                    t.append(time.perf_counter() - t0)
                return min(t) * 1e3

            r = {"image": name, "data": what, "bytes": img_bytes, "n_gpus": n,
                 "host_ms": best(host_gather), "peer_ms": best(peer_gather)}
            if what == "rgb8":
                r["peer_p3_ms"] = best(peer_p3)
                r["host_p3_ms"] = best(host_p3)
                assert peer_p3() == host_p3()
            rows.append(r)
            print(json.dumps(r), flush=True)
            ctx0.close()
            lib.rtclj_host_free(hp)
            lib.rtclj_host_free(tp)
            del dev, d_text
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"gather_n{n}.md")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        f.write(f"| image | data | MB | {n} host copies (ms) | peer copies to GPU 0 + one host copy (ms) | peer + device P3 + text to host (ms) | host gather + host P3 (ms) |\n|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['image']} | {r['data']} | {r['bytes'] / 1e6:.1f} | {r['host_ms']:.2f} | {r['peer_ms']:.2f} | "
                    f"{r.get('peer_p3_ms', float('nan')):.2f} | {r.get('host_p3_ms', float('nan')):.2f} |\n")


if __name__ == "__main__":
    main()
