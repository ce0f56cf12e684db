"""Host-side rows of SURVEY.md 8(f): camera derivation (incl. Clojure's Ratio quirk), the
write-color! quantisation, the P3 encoder, the scene generator -- native C++ against the
Python twins and against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render
from raytracing_clj_b200 import camera as cam_py


def vec(x):
    return (C.c_double * 3)(*x)


def test_ratio_to_double_quirk():
    f = _abi.lib().rtclj_ratio_to_double
    # 16/9 through Ratio.doubleValue is ONE ULP ABOVE 16.0/9.0
    assert f(16, 9).hex() == "0x1.c71c71c71c71dp+0" and (16.0 / 9.0).hex() == "0x1.c71c71c71c71cp+0"
    for num, den in [(16, 9), (400, 225), (1920, 1080), (25, 14), (1, 3), (2, 3), (-7, 3), (400, 224), (10, 4),
                     (123456789, 1000), (1, 1000000007), (999999999999, 7)]:
        assert f(num, den) == cam_py.ratio_to_double(num, den), (num, den)
    assert f(400, 200) == 2.0 and f(0, 5) == 0.0


def test_native_cameras_equal_python_cameras_bit_for_bit():
    lib = _abi.lib()
    for w in (400, 1920, 3840, 37):
        for kw in ({}, R.scenes.COVER_CAMERA, R.scenes.FIELD_CAMERA):
            py = cam_py.main_camera(w, **kw)
            c = _abi.Camera()
            lf, la, vu = py.center, kw.get("look_at", (0.0, 0.0, -1.0)), (0.0, 1.0, 0.0)
            assert lib.rtclj_camera_main(w, py.height, kw.get("vfov", 20.0), vec(lf), vec(la), vec(vu),
                                         kw.get("defocus_angle", 10.0), kw.get("focus_dist", 3.4), C.byref(c)) == 0
            for name in ("pixel00", "pixel_du", "pixel_dv", "center", "defocus_u", "defocus_v"):
                assert tuple(getattr(c, name)) == tuple(getattr(py, name)), (w, name)
        py = cam_py.realm_camera(w)
        c = _abi.Camera()
        assert lib.rtclj_camera_realm(w, py.height, 20.0, vec((-2, 2, 1)), vec((0, 0, -1)), vec((0, 1, 0)), C.byref(c)) == 0
        for name in ("pixel00", "pixel_du", "pixel_dv", "center"):
            assert tuple(getattr(c, name)) == tuple(getattr(py, name)), (w, name)
        py = cam_py.i_camera(w)
        assert lib.rtclj_camera_i(w, py.height, C.byref(c)) == 0
        for name in ("pixel00", "pixel_du", "pixel_dv", "center"):
            assert tuple(getattr(c, name)) == tuple(getattr(py, name)), (w, name)


def test_camera_geometry_sanity():
    c = cam_py.main_camera()
    # the centre of the image looks at look-at: pixel (W/2, H/2) lies on the focus plane along -w
    p = np.array(c.pixel00) + np.array(c.pixel_du) * (c.width / 2 - 0.5) + np.array(c.pixel_dv) * (c.height / 2 - 0.5)
    d = p - np.array(c.center)
    to_at = np.array([0.0, 0.0, -1.0]) - np.array(c.center)
    assert np.allclose(d / np.linalg.norm(d), to_at / np.linalg.norm(to_at), atol=1e-12)
    assert abs(np.linalg.norm(d) - 3.4) < 1e-12  # focus-dist
    assert abs(np.linalg.norm(c.defocus_u) - 3.4 * np.tan(np.radians(5.0))) < 1e-12


def test_quantise_matches_oracle_and_reference_rule():
    # gamma rule (write-color!, raytracing.clj:19-26): any input, incl. negatives, > 1, NaN, inf
    vals = np.concatenate([np.linspace(-0.5, 1.5, 4001), [0.0, 1e-300, 0.999 ** 2, 1.0, np.nan, np.inf, 0.25]])
    got = render.quantise_rgb8(vals, 0)
    want = np.array([O.lib().rto_quantise(float(v), 0) for v in vals], dtype=np.uint8)
    assert np.array_equal(got, want)
    assert render.quantise_rgb8(np.array([0.25, 1.0, 5.0, -1.0, np.nan]), 0).tolist() == [128, 255, 255, 0, 0]
    # linear rule (raytracing_i.clj:170): int(255.999*c), defined for the c in [0,1] that variant produces
    vals = np.concatenate([np.linspace(0.0, 1.0, 4001), [np.nan]])
    got = render.quantise_rgb8(vals, _abi.F_QUANT_LINEAR)
    want = np.array([O.lib().rto_quantise(float(v), 1) for v in vals], dtype=np.uint8)
    assert np.array_equal(got, want) and got[-2] == 255 and got[2000] == 127


def test_ppm_encoder_matches_reference_format():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(5, 7, 3), dtype=np.uint8)
    img[0, 0] = (0, 9, 10)
    img[0, 1] = (99, 100, 255)
    text = render.encode_ppm(img).decode()
    lines = text.split("\n")
    assert lines[:3] == ["P3", "7 5", "255"] and lines[-1] == ""          # raytracing.clj:173
    assert lines[3] == "0 9 10" and lines[4] == "99 100 255"               # write-color!: "r g b\n"
    assert len(lines) == 3 + 35 + 1
    back = np.array(" ".join(lines[3:]).split(), dtype=np.uint8).reshape(5, 7, 3)
    assert np.array_equal(back, img)
    # too-small buffer is reported, not overrun
    n = C.c_size_t()
    buf = C.create_string_buffer(8)
    assert _abi.lib().rtclj_encode_ppm_p3(img.ctypes.data, 7, 5, buf, 8, C.byref(n)) == _abi.E_BUFFER and n.value > 8


def test_ppm_encoder_exact_capacity_and_one_pass_paths_agree():
    """The host writer has a measuring path (tight buffer) and a one-pass path (worst-case buffer)."""
    rng = np.random.default_rng(2)
    for shape in ((1, 1), (1, 2), (3, 5), (64, 33)):
        img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8)
        want = ("P3\n%d %d\n255\n" % (shape[1], shape[0]) + "".join("%d %d %d\n" % tuple(px) for px in img.reshape(-1, 3).tolist())).encode()
        n = C.c_size_t()
        for cap in (len(want), len(want) + 1, len(want) + 3, 64 + 12 * shape[0] * shape[1]):
            buf = C.create_string_buffer(b"\xaa" * (cap + 8), cap + 8)
            assert _abi.lib().rtclj_encode_ppm_p3(img.ctypes.data, shape[1], shape[0], buf, cap, C.byref(n)) == 0
            assert n.value == len(want) and buf.raw[: n.value] == want
            assert buf.raw[cap:] == b"\xaa" * 8                      # nothing past the stated capacity


def test_device_ppm_writer_formatting_primitives(tmp_path):
    """The device P3 writer's byte-parallel digit / length / packing helpers, checked on the CPU:
    exhaustively over value pairs, and against sprintf for random threads' worth of pixels."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "p3_swar_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(root, "raytracing-clj_b200", "csrc"),
                    "-o", exe, os.path.join(root, "tests", "native", "p3_swar_check.cpp")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "ok"


def test_ppm_of_reference_golden_round_trips():
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_images.npz"))["scene_main"]
    text = render.encode_ppm(gold).decode().split("\n")
    assert text[1] == "400 225" and text[3] == "175 198 0"   # first pixel of the reference's scene.ppm
    assert len(text) == 90003 + 1                              # 90 003 lines


def test_ppm_reader_round_trip_and_rejections(tmp_path):
    """The reader half of ppm->png (ppm2png.clj:35-87): our own P3 text back to the pixels, the
    reference's committed render, and the malformed inputs the reference rejects."""
    import io
    from PIL import Image
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, size=(9, 13, 3), dtype=np.uint8)
    assert np.array_equal(render.decode_ppm(render.encode_ppm(img)), img)
    assert np.array_equal(render.decode_ppm(b"P3 2 1 255 1 2 3\t4 5 6"), np.array([[[1, 2, 3], [4, 5, 6]]], dtype=np.uint8))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_images.npz"))["scene_main"]
    assert np.array_equal(render.decode_ppm(render.encode_ppm(gold)), gold)
    for bad in (b"P6\n1 1\n255\n0 0 0\n", b"P3\n1\n", b"P3\n1 1\n256\n0 0 0\n", b"P3\n1 1\n255\n0 0\n",
                b"P3\n1 1\n255\n0 0 0 0\n", b"P3\n1 1\n255\n0 0 300\n", b"P3\n1 1\n255\n0 x 0\n", b"P3\n0 1\n255\n",
                b"P3\n1 1\n100\n0 0 101\n"):
        with pytest.raises(_abi.RtcljError):
            render.decode_ppm(bad)
    src, dst = str(tmp_path / "scene.ppm"), str(tmp_path / "scene.png")
    render.write_ppm(src, gold)
    render.ppm_to_png(src, dst)                                   # (ppm->png "scene.ppm" "scene.png"), raytracing.clj:176
    assert np.array_equal(np.asarray(Image.open(dst).convert("RGB")), gold)


def test_large_images_are_written_by_several_threads_byte_for_byte_like_small_ones():
    """rtclj_encode_ppm_p3 writes images of >= 2^18 pixels with several threads (measure, then write at offsets):
    the text must be the bytes the one-thread loop writes -- checked against Python's own formatting, with exact
    and with generous capacity."""
    import ctypes as C
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, size=(700, 900, 3), dtype=np.uint8)
    img[:100] = 7; img[100:200] = 200; img[-1, -3:] = [[0, 0, 0], [255, 255, 255], [9, 99, 100]]
    want = b"P3\n900 700\n255\n" + b"".join(b"%d %d %d\n" % (r, g, b) for r, g, b in img.reshape(-1, 3).tolist())
    assert render.encode_ppm(img) == want
    lib = _abi.lib()
    n = C.c_size_t()
    exact = C.create_string_buffer(len(want))                       # exactly as many bytes as the text needs
    _abi.check(lib.rtclj_encode_ppm_p3(img.ctypes.data, 900, 700, exact, len(want), C.byref(n)))
    assert n.value == len(want) and exact.raw == want
    short = C.create_string_buffer(len(want) - 1)
    assert lib.rtclj_encode_ppm_p3(img.ctypes.data, 900, 700, short, len(want) - 1, C.byref(n)) == _abi.E_BUFFER and n.value == len(want)


def test_large_ppm_bodies_are_parsed_by_several_threads_with_the_same_results_and_errors():
    """Bodies of >= 4 MB (a 4K scene.ppm has 99 MB) go through the threaded parser; anything unusual falls back
    to the sequential one, so results and rejections are those of small inputs."""
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, size=(900, 1200, 3), dtype=np.uint8)
    img[:300] = 0; img[300:600] = 255                       # one-digit and three-digit stretches: uneven token density
    text = render.encode_ppm(img)
    assert len(text) > (4 << 20)
    assert np.array_equal(render.decode_ppm(text), img)
    assert np.array_equal(render.decode_ppm(text.replace(b"\n", b"\r\n  ").replace(b" ", b"\t ")), img)   # other white space
    body = text.index(b"255\n") + 4
    mid = body + (len(text) - body) // 2
    mid = text.index(b"\n", mid) + 1
    for bad in (text + b"7\n",                                          # one value too many
                text[: text.rindex(b" ")] + b"\n",                      # one value missing
                text[:mid] + b"300 " + text[text.index(b" ", mid) + 1:],  # a value above the maximum, deep in the body
                text[:mid] + b"x" + text[mid + 1:],                      # not a digit
                text[:mid] + b"1.5 " + text[text.index(b" ", mid) + 1:]):
        with pytest.raises(_abi.RtcljError):
            render.decode_ppm(bad)


def test_png_encoder_decodes_to_the_same_pixels():
    import io
    from PIL import Image
    rng = np.random.default_rng(1)
    for shape in ((5, 7, 3), (1, 1, 3), (225, 400, 3), (300, 30000 // 3, 3)):  # the last one spans several 64 KiB blocks
        img = rng.integers(0, 256, size=shape, dtype=np.uint8)
        png = render.encode_png(img)
        assert png[:8] == b"\x89PNG\r\n\x1a\n"
        back = np.asarray(Image.open(io.BytesIO(png)).convert("RGB"))
        assert back.shape == img.shape and np.array_equal(back, img)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_images.npz"))["scene_main"]
    back = np.asarray(Image.open(io.BytesIO(render.encode_png(gold))))
    assert np.array_equal(back, gold)  # like scene.png vs scene.ppm in the reference
    n = C.c_size_t()
    assert _abi.lib().rtclj_encode_png(gold.ctypes.data, 400, 225, (C.c_uint8 * 16)(), 16, C.byref(n)) == _abi.E_BUFFER


def test_native_scene_generator_equals_python_generator():
    lib = _abi.lib()
    for seed, lo, hi in ((7, -11, 11), (3, -11, 11), (7, -50, 50), (1, 0, 0)):
        n = C.c_int32()
        assert lib.rtclj_scene_random_field(seed, lo, hi, 0, None, None, None, None, None, None, C.byref(n)) == 0
        py = R.scenes.to_soa(R.scenes._random_field(seed, lo, hi))
        assert n.value == len(py[1])
        arrs = [np.zeros_like(a) for a in py]
        assert lib.rtclj_scene_random_field(seed, lo, hi, n.value, *(a.ctypes.data for a in arrs), C.byref(n)) == 0
        for a, b in zip(arrs, py):
            assert np.array_equal(a, b)
        assert lib.rtclj_scene_random_field(seed, lo, hi, 2, *(a.ctypes.data for a in arrs), C.byref(n)) in (0, _abi.E_BUFFER)
    cover = R.scenes.cover_hittables(7)
    assert len(cover) == 484 and cover[0][R.hittable.RADIUS] == 1000.0 and cover[-1][R.material.KIND] == R.material.METAL


def test_reference_scene_literals():
    m, r = R.scenes.main_hittables(), R.scenes.realm_hittables()
    assert [b[R.hittable.CENTER] for b in m] == [(0.0, -100.5, -1.0), (0.0, 0.0, -1.2), (-1.0, 0.0, -1.0),
                                                  (-1.0, 0.0, -1.0), (1.0, 0.0, -1.0)]       # raytracing.clj:63-78
    assert [b[R.hittable.RADIUS] for b in m] == [100.0, 0.5, 0.5, 0.4, 0.5]
    assert [b[R.material.KIND] for b in m] == [0, 0, 2, 2, 1]
    assert m[3][R.material.IOR] == 1.00 / 1.5 and m[4][R.material.FUZZ] == 1.0
    assert r[0] == m[1] and r[1] == m[0] and r[2:] == m[2:]                                  # realm/raytracing.clj:292-301


def test_shard_rows_partition_the_image():
    for H, count, rows in ((1080, 8, 4), (225, 2, 4), (7, 3, 5), (100, 4, 64), (9, 1, 4)):
        seen = []
        for i in range(count):
            seen += render.shard_rows(H, i, count, rows)
        assert sorted(seen) == list(range(H))
    assert render.shard_rows(20, 1, 2, 4) == [4, 5, 6, 7, 12, 13, 14, 15]


def test_shard_plan_covers_exactly_the_shards_rows():
    """The library's download plan (C++, rtclj_shard_plan) against the Python shard arithmetic: for every
    shard the planned byte runs are exactly the shard's rows -- no byte twice, none missing, none foreign --
    for ragged last tiles, tiles larger than a staging piece, more shards than tiles, and no sharding."""
    from raytracing_clj_b200 import render
    for H, row_bytes, count, rows, cap in ((225, 96, 2, 4, 0), (135, 64, 8, 1, 0), (37, 1000, 3, 5, 4096),
                                           (50, 4096, 4, 16, 10000), (7, 64, 16, 1, 0), (100, 24, 1, 0, 512),
                                           (2160, 48, 8, 1, 16 << 20), (13, 100, 5, 3, 250)):
        seen = np.zeros(H * row_bytes, dtype=np.int32)
        for idx in range(max(1, count)):
            mine = np.zeros(H * row_bytes, dtype=np.int32)
            for off, pitch, width, height in render.shard_plan(H, row_bytes, idx, count, rows, cap):
                assert 0 < width * height <= (cap or (16 << 20))
                for r in range(height):
                    a = off + r * pitch
                    assert a + width <= H * row_bytes
                    mine[a:a + width] += 1
            want = np.zeros(H * row_bytes, dtype=np.int32)
            for j in render.shard_rows(H, idx, count, rows):
                want[j * row_bytes:(j + 1) * row_bytes] = 1
            assert np.array_equal(mine, want), (H, row_bytes, idx, count, rows)   # each byte of the shard exactly once
            seen += mine
        assert np.all(seen == 1), "the shards partition the image"
