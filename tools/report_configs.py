"""Runs the five BASELINE.json configs on one B200 and writes gpurun_out/r2_configs.{json,md} (copied to profiles/):
device-resident throughput (rays/s = segments/s, samples/s, algorithmic FP32 fraction) and the
parity evidence that fits each size (bit-exact rows against the CPU oracle; statistics against the
reference's committed renders for config 1), plus the reference's `(time ...)` scope end to end: render + P3 text +
PNG + file writes (src/raytracing.clj:99-177).  Usage: python tools/report_configs.py [--quick]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render

QUICK = "--quick" in sys.argv
S, CAM = R.scenes, R.camera
PEAK = 148 * 128 * 2 * 1.965e9 / 1e12


def device_render(world, cam, spp, depth, flags, seed=1, reps=2):
    ctx = render.Context(0)
    ctx.set_scene(world)
    out = torch.zeros((cam.height, cam.width, 3), dtype=torch.float64, device="cuda:0")
    out8 = torch.zeros((cam.height, cam.width, 3), dtype=torch.uint8, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    best = None
    for _ in range(reps):
        ctx.render(cam, spp, depth, seed=seed, flags=flags, d_out_linear=out.data_ptr(), d_out_rgb8=out8.data_ptr(),
                   stream=stream)
        st = ctx.stats(stream)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    ctx.close()
    return out.cpu().numpy(), out8.cpu().numpy(), best


def rows_check(world, cam, spp, depth, flags, seed, lin, rgb, spu, rows):
    """Bit-exact comparison of a few image rows against the CPU oracle."""
    t0 = time.perf_counter()
    bad = 0
    for j in rows:
        lo, ro, _ = O.render(S.to_soa(world), cam, spp, depth, seed=seed, flags=flags, threads=os.cpu_count(),
                             rows=(j, j + 1), samples_per_unit=spu)
        bad += int(not (np.array_equal(lo[j], lin[j]) and np.array_equal(ro[j], rgb[j])))
    return {"rows_checked": list(rows), "rows_differing": bad, "oracle_seconds": round(time.perf_counter() - t0, 1)}


def entry(name, world, cam, spp, depth, flags, st, extra):
    n = len(world)
    sec = st["kernel_ms"] * 1e-3
    e = {"config": name, "n_spheres": n, "image": f"{cam.width}x{cam.height}", "spp": spp, "max_depth": depth,
         "kernel_ms": round(st["kernel_ms"], 2), "rays_per_sec": st["segments"] / sec,
         "samples_per_sec": st["samples"] / sec, "segments_per_sample": st["segments"] / max(1, st["samples"]),
         "fp32_fraction_algorithmic": st["segments"] * (17 * n + 5) / sec / 1e12 / PEAK,
         "fp64_tests_per_segment": st["exact_tests"] / max(1, st["segments"]), "samples_per_unit": st["samples_per_unit"]}
    e.update(extra)
    print(json.dumps(e), flush=True)
    return e


def main_scope():
    """`clojure -M:main` times everything from the camera set-up to ppm->png (src/raytracing.clj:99-177).  The
    same scope here: rtclj_render (host buffers) + P3 text + file write + P3 -> PNG + file write, at the
    reference's default size and at 3840x2160; the P3 text once by the host writer and once by the device
    writer (rtclj_encode_ppm_p3_gpu).  Best of 3, wall clock."""
    import tempfile
    rows = []
    for label, world, cam, spp, depth, flags in (
            ("main default 400x225, 100 spp", S.main_hittables(), CAM.main_camera(), 100, 50, O.FLAGS_MAIN),
            ("main scene 1920x1080, 100 spp", S.main_hittables(), CAM.main_camera(1920), 100, 50, O.FLAGS_MAIN),
            ("raytracing-i 3840x2160, 100 spp", S.i_hittables(), CAM.i_camera(3840), 100, 50, O.FLAGS_I)):
        for writer in ("host", "device"):
            best = None
            for _ in range(3):
                with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
                    t = {}
                    t0 = time.perf_counter()
                    _, rgb8, st = render.render(world, cam, spp, depth, seed=1, flags=flags, want_linear=False)
                    t["render_ms"] = 1e3 * (time.perf_counter() - t0)
                    t1 = time.perf_counter()
                    render.write_ppm(os.path.join(d, "scene.ppm"), rgb8, device=0 if writer == "device" else None)
                    t["ppm_ms"] = 1e3 * (time.perf_counter() - t1)
                    t2 = time.perf_counter()
                    render.ppm_to_png(os.path.join(d, "scene.ppm"), os.path.join(d, "scene.png"))
                    t["png_ms"] = 1e3 * (time.perf_counter() - t2)
                    t["total_ms"] = 1e3 * (time.perf_counter() - t0)
                    t["kernel_ms"] = st["kernel_ms"]
                if best is None or t["total_ms"] < best["total_ms"]:
                    best = t
            rows.append({"scope": label, "p3_writer": writer, **{k: round(v, 2) for k, v in best.items()}})
            print(json.dumps(rows[-1]), flush=True)
    return rows


def main():
    out = []
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_images.npz"))
    # ---- config 1: the reference's default scenes at their default sizes
    for variant, world, cam, flags, g in (("main", S.main_hittables(), CAM.main_camera(), O.FLAGS_MAIN, gold["scene_main"]),
                                          ("realm", S.realm_hittables(), CAM.realm_camera(), O.FLAGS_REALM, gold["scene_realm"])):
        lin, rgb, st = device_render(world, cam, 100, 50, flags)
        lo, ro, so = O.render(S.to_soa(world), cam, 100, 50, seed=1, flags=flags, threads=os.cpu_count(),
                              samples_per_unit=st["samples_per_unit"])
        extra = {"bit_exact_vs_oracle": bool(np.array_equal(lo, lin) and np.array_equal(ro, rgb)),
                 "mean_rgb8": [round(float(x), 3) for x in rgb.reshape(-1, 3).mean(0)],
                 "reference_mean_rgb8": [round(float(x), 3) for x in g.reshape(-1, 3).mean(0)],
                 "mae_vs_reference_render": round(float(np.abs(rgb.astype(int) - g.astype(int)).mean()), 3)}
        out.append(entry(f"1 ({variant} default scene)", world, cam, 100, 50, flags, st, extra))
    # ---- config 2: 5-body material scene, 1920x1080, 100 spp
    world, cam = S.main_hittables(), CAM.main_camera(1920)
    lin, rgb, st = device_render(world, cam, 100, 50, O.FLAGS_MAIN)
    out.append(entry("2 (material scene 1080p)", world, cam, 100, 50, O.FLAGS_MAIN, st,
                     rows_check(world, cam, 100, 50, O.FLAGS_MAIN, 1, lin, rgb, st["samples_per_unit"], (3, 540, 1000))))
    # ---- config 3: cover scene, 1920x1080, 500 spp (the bench workload)
    world, cam = S.cover_hittables(7), CAM.main_camera(1920, 1080, **S.COVER_CAMERA)
    spp = 50 if QUICK else 500
    lin, rgb, st = device_render(world, cam, spp, 50, O.FLAGS_MAIN, reps=1 if QUICK else 2)
    out.append(entry("3 (RTIOW cover scene 1080p)", world, cam, spp, 50, O.FLAGS_MAIN, st,
                     rows_check(world, cam, spp, 50, O.FLAGS_MAIN, 1, lin, rgb, st["samples_per_unit"], (700,) if not QUICK else (700, 900))))
    # ---- config 4: primary-ray renders at 3840x2160, 8-bit bit-exact (whole image)
    cam = CAM.i_camera(3840)
    for label, world, flags, depth in (("4i (raytracing-i normal shading 4K)", S.i_hittables(), O.FLAGS_I, 50),
                                       ("4ii (realm, max-depth 1, camera sees sky, 4K)", S.realm_hittables(), O.FLAGS_REALM, 1)):
        spp = 16 if QUICK else 100
        lin, rgb, st = device_render(world, cam, spp, depth, flags)
        lo, ro, so = O.render(S.to_soa(world), cam, spp, depth, seed=1, flags=flags, threads=os.cpu_count(),
                              samples_per_unit=st["samples_per_unit"])
        out.append(entry(label, world, cam, spp, depth, flags, st,
                         {"rgb8_bit_exact_whole_image": bool(np.array_equal(ro, rgb)),
                          "linear_bit_exact_whole_image": bool(np.array_equal(lo, lin))}))
    # ---- config 5: ~10 000 spheres, 3840x2160, 256 spp, brute force
    world, cam = S.field_hittables(7), CAM.main_camera(3840, 2160, **S.FIELD_CAMERA)
    spp = 8 if QUICK else 256
    lin, rgb, st = device_render(world, cam, spp, 50, O.FLAGS_MAIN, reps=1)
    # parity at this size: one row of an 8-spp render of the same scene/camera against the oracle
    lin8, rgb8, st8 = device_render(world, cam, 8, 50, O.FLAGS_MAIN, reps=1)
    chk = rows_check(world, cam, 8, 50, O.FLAGS_MAIN, 1, lin8, rgb8, st8["samples_per_unit"], (1500,))
    chk["rows_check_spp"] = 8
    out.append(entry("5 (10k-sphere field 4K)", world, cam, spp, 50, O.FLAGS_MAIN, st, chk))
    scope = main_scope()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_configs.json"), "w") as f:
        json.dump({"configs": out, "main_scope": scope}, f, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "r2_configs.md"), "w") as f:
        f.write("| config | spheres | image | spp | kernel ms | rays/s | samples/s | seg/sample | FP32 fraction (algorithmic) | parity |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for e in out:
            par = {k: v for k, v in e.items() if "exact" in k or "rows_" in k or "mae" in k}
            f.write(f"| {e['config']} | {e['n_spheres']} | {e['image']} | {e['spp']} | {e['kernel_ms']} | {e['rays_per_sec']:.3e} | "
                    f"{e['samples_per_sec']:.3e} | {e['segments_per_sample']:.3f} | {100*e['fp32_fraction_algorithmic']:.1f} % | {par} |\n")
        f.write("\nThe reference's `(time ...)` scope (render + P3 + file + PNG + file), best of 3, wall clock, files on /dev/shm:\n\n"
                "| scope | P3 writer | render ms (kernel ms) | P3 + write ms | P3 -> PNG + write ms | total ms |\n|---|---|---|---|---|---|\n")
        for r in scope:
            f.write(f"| {r['scope']} | {r['p3_writer']} | {r['render_ms']} ({r['kernel_ms']}) | {r['ppm_ms']} | {r['png_ms']} | {r['total_ms']} |\n")


if __name__ == "__main__":
    main()
