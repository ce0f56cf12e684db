"""ctypes wrapper around oracle/_build/librt_oracle.so -- the CHECKER, used only by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "librt_oracle.so")

F_NEAR_ZERO_GUARD, F_SCHLICK, F_REVERSE_PRODUCT, F_MEAN_DIVIDE, F_NORMAL_SHADING, F_QUANT_LINEAR = 1, 2, 4, 8, 16, 32
FLAGS_MAIN = F_NEAR_ZERO_GUARD | F_SCHLICK | F_REVERSE_PRODUCT | F_MEAN_DIVIDE
FLAGS_REALM = 0
FLAGS_I = F_NORMAL_SHADING | F_QUANT_LINEAR


class Scene(C.Structure):
    _fields_ = [("n", C.c_int32), ("_pad", C.c_int32), ("center_xyz", C.c_void_p),
                ("radius", C.c_void_p), ("material", C.c_void_p), ("albedo_rgb", C.c_void_p),
                ("fuzz", C.c_void_p), ("ior", C.c_void_p)]


class Camera(C.Structure):
    _fields_ = [("pixel00", C.c_double * 3), ("pixel_du", C.c_double * 3),
                ("pixel_dv", C.c_double * 3), ("center", C.c_double * 3),
                ("defocus_u", C.c_double * 3), ("defocus_v", C.c_double * 3),
                ("defocus_angle", C.c_double), ("width", C.c_int32), ("height", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("spp", C.c_int32), ("max_depth", C.c_int32), ("seed", C.c_uint64),
                ("flags", C.c_uint32), ("samples_per_unit", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("segments", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("rng_blocks", C.c_uint64), ("hits", C.c_uint64 * 3), ("seg_hist", C.c_uint64 * 64)]


_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("rt_oracle.c", "rt_oracle.h", "Makefile")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.rto_render.restype = C.c_int
        _lib.rto_render.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.POINTER(Params), C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        _lib.rto_render_strided.restype = C.c_int
        _lib.rto_render_strided.argtypes = [C.POINTER(Scene), C.POINTER(Camera), C.POINTER(Params), C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        _lib.rto_uniform.restype = C.c_double
        _lib.rto_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        _lib.rto_philox4x32_10.restype = None
        _lib.rto_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        _lib.rto_hit_anything.restype = C.c_int
        _lib.rto_hit_anything.argtypes = [C.POINTER(Scene), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.c_double, C.c_double, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _lib.rto_quantise.restype = C.c_int
        _lib.rto_quantise.argtypes = [C.c_double, C.c_int]
    return _lib


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    o = (C.c_uint32 * 4)()
    lib().rto_philox4x32_10(c, key[0], key[1], o)
    return tuple(int(x) for x in o)


def make_scene(soa):
    center, radius, kind, albedo, fuzz, ior = soa
    s = Scene(len(radius), 0, center.ctypes.data, radius.ctypes.data, kind.ctypes.data,
              albedo.ctypes.data, fuzz.ctypes.data, ior.ctypes.data)
    s._keep = soa
    return s


def make_camera(cam) -> Camera:
    c = Camera()
    for name in ("pixel00", "pixel_du", "pixel_dv", "center", "defocus_u", "defocus_v"):
        getattr(c, name)[:] = [float(x) for x in getattr(cam, name)]
    c.defocus_angle = float(cam.defocus_angle)
    c.width, c.height = int(cam.width), int(cam.height)
    return c


def render(soa, cam, spp, max_depth, seed=1, flags=FLAGS_MAIN, threads=1, rows=None,
           samples_per_unit=0, want_rgb8=True, row_step=1):
    """Returns (linear float64 [H,W,3], rgb8 uint8 [H,W,3] or None, Stats)."""
    sc, cm = make_scene(soa), make_camera(cam)
    prm = Params(int(spp), int(max_depth), int(seed), int(flags), int(samples_per_unit))
    H, W = cm.height, cm.width
    lin = np.zeros((H, W, 3), dtype=np.float64)
    rgb = np.zeros((H, W, 3), dtype=np.uint8) if want_rgb8 else None
    st = Stats()
    r0, r1 = (0, H) if rows is None else rows
    rc = lib().rto_render_strided(C.byref(sc), C.byref(cm), C.byref(prm), int(threads), int(r0), int(r1),
                                  int(row_step), lin.ctypes.data, rgb.ctypes.data if want_rgb8 else None,
                                  C.byref(st))
    if rc != 0:
        raise RuntimeError(f"rto_render failed: {rc}")
    return lin, rgb, st


def hit_anything(soa, origin, direction, t_min=1e-3, t_max=float("inf")):
    sc = make_scene(soa)
    o = (C.c_double * 3)(*origin)
    d = (C.c_double * 3)(*direction)
    t = C.c_double()
    p = (C.c_double * 3)()
    n = (C.c_double * 3)()
    ff = C.c_int()
    idx = lib().rto_hit_anything(C.byref(sc), o, d, t_min, t_max, C.byref(t), p, n, C.byref(ff))
    return idx, t.value, tuple(p), tuple(n), bool(ff.value)
