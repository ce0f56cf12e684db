"""Row f-1 of SURVEY.md section 8: the P3 writer (src/raytracing.clj:172-175) as device kernels.
Byte work => bit-exact against the host writer and against a plain-Python formatting of the
reference's `write-color!` lines, through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_images.npz")


def python_p3(img):
    H, W, _ = img.shape
    return (f"P3\n{W} {H}\n255\n" + "".join(f"{r} {g} {b}\n" for r, g, b in img.reshape(-1, 3).tolist())).encode()


@pytest.mark.parametrize("shape", [(1, 1), (1, 3), (5, 7), (1, 1023), (1, 1024), (1, 1025), (3, 1024), (33, 97), (225, 400)])
def test_device_writer_matches_python_formatting(shape):
    rng = np.random.default_rng(shape[0] * 100003 + shape[1])
    img = rng.integers(0, 256, (shape[0], shape[1], 3), dtype=np.uint8)
    assert render.encode_ppm(img, device=0) == python_p3(img)


@pytest.mark.parametrize("fill", [0, 9, 10, 99, 100, 255])
def test_digit_count_boundaries(fill):
    img = np.full((17, 129, 3), fill, dtype=np.uint8)
    img[3, 5] = (0, 10, 100)
    img[16, 128] = (255, 9, 99)
    assert render.encode_ppm(img, device=0) == python_p3(img)


def test_reference_renders_round_trip():
    gold = np.load(GOLD)
    for key in ("scene_main", "scene_realm"):
        img = gold[key]
        text = render.encode_ppm(img, device=0)
        assert text == render.encode_ppm(img)
        body = np.array(text.split()[4:], dtype=np.int64).reshape(img.shape)
        assert np.array_equal(body, img)


def test_full_size_images_match_host_writer():
    import torch
    rng = np.random.default_rng(11)
    ctx = render.Context(0)
    stream = torch.cuda.current_stream().cuda_stream
    for H, W in ((1080, 1920), (2160, 3840)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        img[: H // 3] //= 26   # a band of one-digit values, a band of two-digit ones
        img[H // 3: H // 2] //= 3
        want = render.encode_ppm(img)
        assert render.encode_ppm(img, device=0) == want
        d_img = torch.from_numpy(img).cuda()
        for cap in (len(want), 64 + 12 * W * H):   # exact-size and worst-case buffers: the two writer paths
            d_txt = torch.zeros(cap, dtype=torch.uint8, device="cuda:0")
            for _ in range(3):                     # the look-back must not depend on scheduling luck
                n = ctx.encode_ppm(d_img.data_ptr(), W, H, d_txt.data_ptr(), cap, stream)
                assert n == len(want) and bytes(d_txt[:n].cpu().numpy()) == want
    ctx.close()


def test_device_pointers_any_alignment_and_capacity_errors():
    import torch
    ctx = render.Context(0)
    rng = np.random.default_rng(5)
    W, H = 61, 43
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    want = python_p3(img)
    stream = torch.cuda.current_stream().cuda_stream
    for in_off in (0, 1, 2, 3):
        for out_off in (0, 1, 7, 15):
            src = torch.zeros(img.size + 8, dtype=torch.uint8, device="cuda:0")
            src[in_off:in_off + img.size] = torch.from_numpy(img.reshape(-1)).cuda()
            dst = torch.full((len(want) + 64,), 0xAA, dtype=torch.uint8, device="cuda:0")
            n = ctx.encode_ppm(src.data_ptr() + in_off, W, H, 0, 0, stream)
            assert n == len(want)
            n = ctx.encode_ppm(src.data_ptr() + in_off, W, H, dst.data_ptr() + out_off, len(want), stream)
            got = dst.cpu().numpy()
            assert bytes(got[out_off:out_off + n]) == want
            assert (got[:out_off] == 0xAA).all() and (got[out_off + n:] == 0xAA).all()  # nothing outside the text
            big = torch.full((64 + 12 * W * H + 32,), 0xAA, dtype=torch.uint8, device="cuda:0")
            n = ctx.encode_ppm(src.data_ptr() + in_off, W, H, big.data_ptr() + out_off, 64 + 12 * W * H, stream)
            got = big.cpu().numpy()
            assert bytes(got[out_off:out_off + n]) == want
            assert (got[:out_off] == 0xAA).all() and (got[out_off + n:] == 0xAA).all()
    src = torch.from_numpy(img.reshape(-1)).cuda()
    dst = torch.zeros(16 * W * H, dtype=torch.uint8, device="cuda:0")
    # both writer paths: exact-size buffer (count + scan + write) and worst-case buffer (single pass)
    for cap in (len(want), 64 + 12 * W * H):
        n = ctx.encode_ppm(src.data_ptr(), W, H, dst.data_ptr(), cap, stream)
        assert bytes(dst[:n].cpu().numpy()) == want
    with pytest.raises(_abi.RtcljError) as e:
        ctx.encode_ppm(src.data_ptr(), W, H, dst.data_ptr(), len(want) - 1, stream)
    assert e.value.code == _abi.E_BUFFER
    assert min(ctx.encode_ms()) >= 0.0
    ctx.close()
    n = C.c_size_t()
    assert _abi.lib().rtclj_encode_ppm_p3_gpu(0, None, 4, 4, None, 0, C.byref(n)) == _abi.E_INVALID
    buf = C.create_string_buffer(8)
    small = np.zeros((5, 7, 3), dtype=np.uint8)
    assert _abi.lib().rtclj_encode_ppm_p3_gpu(0, small.ctypes.data, 7, 5, buf, 8, C.byref(n)) == _abi.E_BUFFER and n.value > 8


def test_render_then_encode_on_device():
    """The reference's tail: render, then write-color! per pixel -- both steps on the device."""
    import torch
    world, cam = R.scenes.main_hittables(), R.camera.main_camera()
    ctx = render.Context(0)
    ctx.set_scene(world)
    rgb = torch.zeros((cam.height, cam.width, 3), dtype=torch.uint8, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    ctx.render(cam, 4, 50, seed=1, d_out_rgb8=rgb.data_ptr(), stream=stream)
    cap = 64 + 12 * cam.width * cam.height
    text = torch.zeros(cap, dtype=torch.uint8, device="cuda:0")
    n = ctx.encode_ppm(rgb.data_ptr(), cam.width, cam.height, text.data_ptr(), cap, stream)
    ctx.close()
    assert bytes(text[:n].cpu().numpy()) == python_p3(rgb.cpu().numpy())
