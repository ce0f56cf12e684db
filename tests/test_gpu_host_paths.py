"""GPU tests of the host-facing paths added in round 2: one process driving
several GPUs must OVERLAP them, renders of different scenes may run concurrently on one device, primary-ray
renders default to the reference's strict summation order, and pinned caller buffers are written directly."""
import ctypes as C
import os
import time

import numpy as np
import pytest

import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render

pytestmark = pytest.mark.gpu
S, CAM = R.scenes, R.camera


def device_count():
    n = C.c_int()
    _abi.check(_abi.lib().rtclj_device_count(C.byref(n)))
    return n.value


def test_multi_device_equals_single_device_and_overlaps():
    """rtclj_render_multi (src/raytracing.clj:157-171 -- the reference's pool -- as one process over N GPUs):
    same image as one GPU, and N devices must take <= 1.15 x (one-device time / N): they render and copy
    at the same time instead of one after the other (round 1 serialised them)."""
    n = device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    world, cam = S.cover_hittables(7), CAM.main_camera(1920, 1080, **S.COVER_CAMERA)
    soa = S.to_soa(world)
    rgb1 = np.zeros((cam.height, cam.width, 3), dtype=np.uint8)
    rgbn = np.zeros_like(rgb1)

    def wall(devices, out):
        best, st = 1e9, None
        for _ in range(3):
            t0 = time.perf_counter()
            _, _, st = render.render(soa, cam, 96, 50, seed=2, devices=devices, want_linear=False, out_rgb8=out)
            best = min(best, time.perf_counter() - t0)
        return best, st

    wall([0], rgb1)  # warm-up: contexts, staging buffers
    wall(list(range(n)), rgbn)
    t1, st1 = wall([0], rgb1)
    tn, stn = wall(list(range(n)), rgbn)
    assert np.array_equal(rgb1, rgbn) and st1["segments"] == stn["segments"] and stn["n_devices"] == n
    assert tn <= 1.15 * t1 / n, f"{n} devices took {tn * 1e3:.1f} ms, one device {t1 * 1e3:.1f} ms"
    # the linear image too, into pageable memory (staged download), any tile height
    a, ra, sa = render.render(soa, CAM.main_camera(192, 108, **S.COVER_CAMERA), 8, 50, seed=2, devices=[0])
    for rows in (0, 1, 4, 7):
        prm_rows = dict(devices=list(range(n)))
        b, rb, sb = render.render(soa, CAM.main_camera(192, 108, **S.COVER_CAMERA), 8, 50, seed=2, **prm_rows)
        assert np.array_equal(a, b) and np.array_equal(ra, rb) and sa["segments"] == sb["segments"], rows


def test_device_list_is_validated():
    for bad in ([-1], [0, 0], [device_count()]):
        with pytest.raises(_abi.RtcljError) as e:
            render.render(S.main_hittables(), CAM.main_camera(16), 1, 5, devices=bad)
        assert e.value.code == _abi.E_INVALID and str(e.value)


@pytest.mark.parametrize("which", ["lane", "lane2", "wave", "split"])
def test_two_scenes_render_concurrently_on_one_device(which):
    """Two contexts, two streams, two different small scenes, launches interleaved without any
    synchronisation between them.  The cull table travels with each launch (kernel parameters); with
    round 1's module-global __constant__ table the second upload corrupted the first render."""
    import torch
    extra = {"lane": _abi.F_LANE_KERNEL, "lane2": _abi.F_LANE2_KERNEL, "wave": _abi.F_WAVE_KERNEL,
             "split": _abi.F_SPLIT_KERNEL}[which]
    jobs = [(S.cover_hittables(7), CAM.main_camera(256, 144, **S.COVER_CAMERA), _abi.FLAGS_MAIN, 3),
            (S.realm_hittables(), CAM.realm_camera(256), _abi.FLAGS_REALM, 4)]
    want = [render.render(w, cam, 24, 50, seed=seed, flags=fl | extra, samples_per_unit=8) for w, cam, fl, seed in jobs]
    ctxs, outs, streams = [], [], []
    for w, cam, fl, seed in jobs:
        c = render.Context(0)
        c.set_scene(w)
        ctxs.append(c)
        outs.append((torch.zeros((cam.height, cam.width, 3), dtype=torch.float64, device="cuda:0"),
                     torch.zeros((cam.height, cam.width, 3), dtype=torch.uint8, device="cuda:0")))
        streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()
    for _ in range(4):  # back to back: the second launch is issued while the first kernel runs
        for (w, cam, fl, seed), c, (lin, rgb), s in zip(jobs, ctxs, outs, streams):
            c.render(cam, 24, 50, seed=seed, flags=fl | extra, samples_per_unit=8, d_out_linear=lin.data_ptr(),
                     d_out_rgb8=rgb.data_ptr(), stream=s.cuda_stream)
    torch.cuda.synchronize()
    for (lin_w, rgb_w, st_w), c, (lin, rgb), s in zip(want, ctxs, outs, streams):
        assert c.stats(s.cuda_stream)["segments"] == st_w["segments"]
        assert np.array_equal(lin.cpu().numpy(), lin_w) and np.array_equal(rgb.cpu().numpy(), rgb_w)
        c.close()


def test_primary_ray_renders_default_to_the_strict_order_4k_100spp():
    """BASELINE.json config 4 at its full size AND sample count (3840x2160, 100 spp): primary-ray renders
    (normal shading, or max-depth 1) default to one sequential sum per pixel -- the reference's order,
    src/experimental/raytracing_i.clj:146-163 -- and equal the oracle bit for bit; the chunked mode gives the
    same 8-bit image."""
    cam = CAM.i_camera(3840)
    world = S.i_hittables()
    lin_s, rgb_s, st_s = render.render(world, cam, 100, 50, seed=4, flags=_abi.FLAGS_I)
    assert st_s["samples_per_unit"] == 100 and st_s["segments"] == 3840 * 2160 * 100
    _, rgb_c, st_c = render.render(world, cam, 100, 50, seed=4, flags=_abi.FLAGS_I, samples_per_unit=20, want_linear=False)
    assert st_c["samples_per_unit"] == 20 and np.array_equal(rgb_s, rgb_c)
    lin_o, rgb_o, st_o = O.render(S.to_soa(world), cam, 100, 50, seed=4, flags=O.FLAGS_I, threads=os.cpu_count() or 8)
    assert np.array_equal(rgb_o, rgb_s) and np.array_equal(lin_o, lin_s) and st_o.segments == st_s["segments"]
    del lin_o, rgb_o, lin_s
    # realm semantics at max-depth 1 on a camera that sees sky: strict by default as well
    _, rgb_d, st_d = render.render(S.realm_hittables(), cam, 12, 1, seed=6, flags=_abi.FLAGS_REALM, want_linear=False)
    assert st_d["samples_per_unit"] == 12
    # a full-depth render keeps the chunked default (load balance); explicit strict is honoured
    _, _, st_f = render.render(S.main_hittables(), CAM.main_camera(64), 64, 50, seed=1, want_linear=False)
    _, _, st_x = render.render(S.main_hittables(), CAM.main_camera(64), 64, 50, seed=1, samples_per_unit=64, want_linear=False)
    assert st_f["samples_per_unit"] < 64 and st_x["samples_per_unit"] == 64


def test_pinned_caller_buffers_are_written_directly_and_equal_pageable_ones():
    lib = _abi.lib()
    world, cam = S.main_hittables(), CAM.main_camera(320)
    H, W = cam.height, cam.width
    lin_p, rgb_p, st_p = render.render(world, cam, 16, 50, seed=9)       # pageable numpy: staged inside the library
    ptr = C.c_void_p()
    _abi.check(lib.rtclj_host_alloc(H * W * 3 * 9, C.byref(ptr)))
    try:
        lin = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(H, W, 3))
        rgb = np.ctypeslib.as_array(C.cast(C.c_void_p(ptr.value + H * W * 24), C.POINTER(C.c_uint8)), shape=(H, W, 3))
        lin[:] = -1.0
        rgb[:] = 7
        _, _, st = render.render(world, cam, 16, 50, seed=9, out_linear=lin, out_rgb8=rgb)
        assert np.array_equal(lin, lin_p) and np.array_equal(rgb, rgb_p) and st["segments"] == st_p["segments"]
        # sharded into pinned memory: only the shard's rows are written
        lin[:] = -1.0
        render.render(world, cam, 16, 50, seed=9, shard=(1, 3, 2), out_linear=lin, out_rgb8=rgb)
        mine = render.shard_rows(H, 1, 3, 2)
        other = [j for j in range(H) if j not in set(mine)]
        assert np.array_equal(lin[mine], lin_p[mine]) and np.all(lin[other] == -1.0)
    finally:
        _abi.check(lib.rtclj_host_free(ptr))
    # registering memory the caller owns
    buf = np.zeros((H, W, 3), dtype=np.float64)
    _abi.check(lib.rtclj_host_register(C.c_void_p(buf.ctypes.data), buf.nbytes))
    try:
        render.render(world, cam, 16, 50, seed=9, out_linear=buf, want_rgb8=False)
        assert np.array_equal(buf, lin_p)
    finally:
        _abi.check(lib.rtclj_host_unregister(C.c_void_p(buf.ctypes.data)))


@pytest.mark.parametrize("extra", [_abi.F_LANE_KERNEL, _abi.F_LANE2_KERNEL, _abi.F_SMEM_TABLE, _abi.F_WAVE_KERNEL,
                                   _abi.F_SPLIT_KERNEL])
def test_strict_order_on_long_paths_goes_through_the_sample_buffer(extra):
    """samples_per_unit >= spp on a full-depth render of >= 64 spp: the samples are traced in small units,
    every sample's colour is stored and finalize_kernel adds them in sample order -- bit for bit the
    reference's sequential sum (src/raytracing.clj:142-155), i.e. the oracle in strict mode.  (The wavefront
    kernel keeps one unit per pixel.)  Also sharded, ragged, and in realm's x (1/spp) form."""
    for world, cam, spp, flags, seed in ((S.main_hittables(), CAM.main_camera(96), 64, O.FLAGS_MAIN, 5),
                                         (S.cover_hittables(7), CAM.main_camera(64, 36, **S.COVER_CAMERA), 70, O.FLAGS_MAIN, 6),
                                         (S.realm_hittables(), CAM.realm_camera(37), 65, O.FLAGS_REALM, 7)):
        soa = S.to_soa(world)
        lin_o, rgb_o, st_o = O.render(soa, cam, spp, 50, seed=seed, flags=flags, threads=8, samples_per_unit=spp)
        lin_g, rgb_g, st_g = render.render(soa, cam, spp, 50, seed=seed, flags=flags | extra, samples_per_unit=spp)
        assert st_g["samples_per_unit"] == spp and st_g["segments"] == st_o.segments
        assert np.array_equal(lin_o, lin_g) and np.array_equal(rgb_o, rgb_g)
        lin = np.zeros_like(lin_g)
        for idx in range(3):
            render.render(soa, cam, spp, 50, seed=seed, flags=flags | extra, samples_per_unit=spp, shard=(idx, 3, 2),
                          out_linear=lin, want_rgb8=False)
        assert np.array_equal(lin, lin_o)


def test_render_to_ppm_text_equals_render_then_encode():
    """rtclj_render_multi_ppm: render loop + write-color! loop (raytracing.clj:141-175) in one call, the shards
    assembled on devices[0] by device-to-device copies and encoded there.  Same bytes as rendering to host
    buffers and running the host P3 writer, for one GPU and for every GPU of the box, any tile height."""
    n = device_count()
    for world, cam, spp, depth, flags in ((S.main_hittables(), CAM.main_camera(200), 8, 50, _abi.FLAGS_MAIN),
                                          (S.i_hittables(), CAM.i_camera(333), 4, 50, _abi.FLAGS_I),
                                          (S.cover_hittables(7), CAM.main_camera(97, 55, **S.COVER_CAMERA), 6, 50, _abi.FLAGS_MAIN)):
        _, rgb, st = render.render(world, cam, spp, depth, seed=3, flags=flags, want_linear=False)
        want = render.encode_ppm(rgb)
        for devs in ([0], list(range(n)), list(range(n))[::-1]):
            text, st2 = render.render_ppm(world, cam, spp, depth, seed=3, flags=flags, devices=devs)
            assert text == want, devs
            assert st2["segments"] == st["segments"] and st2["n_devices"] == len(devs)
    # sizing call and a buffer that is too small
    lib = _abi.lib()
    sc, cm = render._scene_struct(S.to_soa(S.main_hittables())), render._camera_struct(CAM.main_camera(64))
    prm = _abi.Params(2, 5, 1, _abi.FLAGS_MAIN, 0, 0, 0, 0, 0, 0)
    arr = (C.c_int32 * 1)(0)
    ln = C.c_size_t()
    assert lib.rtclj_render_multi_ppm(C.byref(sc), C.byref(cm), C.byref(prm), arr, 1, None, 0, C.byref(ln), None) == 0
    assert ln.value >= 64 * 36 * 6
    small = C.create_string_buffer(100)
    assert lib.rtclj_render_multi_ppm(C.byref(sc), C.byref(cm), C.byref(prm), arr, 1, small, 100, C.byref(ln), None) == _abi.E_BUFFER
    assert ln.value > 100 and b"capacity" in lib.rtclj_last_error()
