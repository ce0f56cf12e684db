"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on
the same Philox stream.  Bar: BIT-EXACT linear radiance (float64), 8-bit output and
segment counts -- stricter than north_star's tolerance (>= 99.9 % of pixels within 1e-3),
which the cull + fp64 design makes unnecessary.  Run with `pytest -m gpu` on a B200."""
import os

import numpy as np
import pytest

import oracle_lib as O
import raytracing_clj_b200 as R
from raytracing_clj_b200 import _abi, render

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
S, CAM = R.scenes, R.camera


def gpu(world, cam, spp, depth, **kw):
    return render.render(world, cam, spp, depth, **kw)


def assert_same(world, cam, spp, depth, seed, flags, unit=0, threads=8, **kw):
    soa = S.to_soa(world)
    lin_o, rgb_o, st_o = O.render(soa, cam, spp, depth, seed=seed, flags=flags & 0xffff, threads=threads,
                                  samples_per_unit=unit if unit else spp)   # the high bits select GPU code paths
    lin_g, rgb_g, st_g = gpu(soa, cam, spp, depth, seed=seed, flags=flags,
                             samples_per_unit=unit if unit else spp, **kw)
    bad = np.argwhere(lin_o != lin_g)
    assert bad.size == 0, (f"{len(bad)} of {lin_o.size} linear values differ; first {bad[0]}: "
                           f"oracle {lin_o[tuple(bad[0])]!r} gpu {lin_g[tuple(bad[0])]!r}")
    assert np.array_equal(rgb_o, rgb_g)
    assert st_g["segments"] == st_o.segments and st_g["samples"] == st_o.samples
    return st_g


def test_library_loaded_and_device_present():
    import ctypes as C
    n = C.c_int()
    _abi.check(_abi.lib().rtclj_device_count(C.byref(n)))
    assert n.value >= 1


def test_committed_fixtures_bit_exact():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fx = np.load(os.path.join(GOLD, "oracle_fixtures.npz"))
    for name, (bodies, cam, spp, depth, seed, flags, unit) in mg.fixture_cases().items():
        lin, rgb, st = gpu(bodies, cam, spp, depth, seed=seed, flags=flags,
                           samples_per_unit=unit if unit else spp)
        assert np.array_equal(lin, fx[name + "/linear"]), name
        assert np.array_equal(rgb, fx[name + "/rgb8"]), name
        assert st["segments"] == int(fx[name + "/segments"][0]), name


@pytest.mark.parametrize("variant", ["main", "realm", "i"])
def test_reference_scenes_bit_exact(variant):
    if variant == "main":
        assert_same(S.main_hittables(), CAM.main_camera(200), 32, 50, 1, O.FLAGS_MAIN)
    elif variant == "realm":
        assert_same(S.realm_hittables(), CAM.realm_camera(200), 32, 50, 2, O.FLAGS_REALM)
    else:
        assert_same(S.i_hittables(), CAM.i_camera(200), 32, 50, 3, O.FLAGS_I)


def test_config1_default_scene_statistics_vs_reference_render():
    """BASELINE.json config 1 on the GPU: 400x225, 100 spp, depth 50 -- bit-exact vs the
    oracle AND statistically consistent with the reference's committed scene.ppm."""
    gold = np.load(os.path.join(GOLD, "reference_images.npz"))["scene_main"]
    st = assert_same(S.main_hittables(), CAM.main_camera(), 100, 50, 1, O.FLAGS_MAIN)
    _, rgb, _ = gpu(S.main_hittables(), CAM.main_camera(), 100, 50, seed=1, flags=O.FLAGS_MAIN,
                    samples_per_unit=100)
    assert np.all(np.abs(rgb.reshape(-1, 3).mean(0) - gold.reshape(-1, 3).mean(0)) < 0.1)
    assert abs(st["segments"] / st["samples"] - 3.675) < 0.01


def test_cover_scene_bit_exact():
    world = S.cover_hittables(7)
    assert 480 <= len(world) <= 488
    st = assert_same(world, CAM.main_camera(160, 90, **S.COVER_CAMERA), 8, 50, 7, O.FLAGS_MAIN)
    # list flushes (a ray whose survivors span more than 16 half blocks) must stay rare
    assert st["list_overflows"] < 1e-4 * st["segments"]
    # the cull must leave only a handful of fp64 tests per segment
    assert st["exact_tests"] / st["segments"] < 12


def test_ten_thousand_sphere_field_bit_exact():
    world = S.field_hittables(7)
    assert 9900 <= len(world) <= 10004
    assert_same(world, CAM.main_camera(96, 54, **S.FIELD_CAMERA), 4, 50, 5, O.FLAGS_MAIN)


def test_cull_equals_exhaustive_fp64_scan():
    world = S.cover_hittables(3)
    cam = CAM.main_camera(128, 72, **S.COVER_CAMERA)
    a, ra, sa = gpu(world, cam, 8, 50, seed=9, flags=O.FLAGS_MAIN, samples_per_unit=8)
    b, rb, sb = gpu(world, cam, 8, 50, seed=9, flags=O.FLAGS_MAIN | _abi.F_NO_CULL, samples_per_unit=8)
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and sa["segments"] == sb["segments"]
    assert sb["exact_tests"] == sb["segments"] * len(world)


@pytest.mark.parametrize("n", [1, 2, 3, 6, 7])
def test_handful_of_spheres_scans_exhaustively_and_stays_bit_exact(n):
    """Scenes of one or two spheres (six for primary-ray renders) skip the fp32 cull: rtclj_ctx_render picks the
    exhaustive fp64 scan by itself (faster there, rtclj_abi.cu).  Same image either way, and equal to the oracle."""
    world = (S.main_hittables() + S.cover_hittables(7)[4:8])[:n]
    for flags, cam, scans in ((O.FLAGS_MAIN, CAM.main_camera(160), n <= 2), (O.FLAGS_I, CAM.i_camera(160), n <= 6)):
        st = assert_same(world, cam, 12, 50, 5, flags)
        assert (st["exact_tests"] == st["segments"] * n) == scans, (n, flags, st)
        lin_c, rgb_c, st_c = gpu(world, cam, 12, 50, seed=5, flags=flags | _abi.F_LANE2_KERNEL, samples_per_unit=12)  # culls
        lin_d, rgb_d, st_d = gpu(world, cam, 12, 50, seed=5, flags=flags, samples_per_unit=12)
        assert np.array_equal(lin_c, lin_d) and np.array_equal(rgb_c, rgb_d) and st_c["segments"] == st_d["segments"]
        assert st_c["exact_tests"] < st_c["segments"] * n or n == 1


@pytest.mark.parametrize("n", [0, 1, 2, 5, 6])
def test_primary_ray_kernel_equals_the_general_kernel_and_the_oracle(n):
    """Primary rays only (normal shading, or max-depth 1) on <= 6 spheres run render_primary_kernel; RTCLJ_F_LANE_KERNEL
    keeps render_kernel.  Same linear image, 8-bit image and counters: whole pixels (strict), chunks, a shard, with
    and without the defocus disk, both quantisations -- and equal to the oracle."""
    world = (S.main_hittables() + S.cover_hittables(7)[4:8])[:n]
    cases = (("normal", O.FLAGS_I, CAM.i_camera(160), 50), ("normal-depth1", O.FLAGS_I, CAM.i_camera(96), 1),
             ("main-depth1-defocus", O.FLAGS_MAIN, CAM.main_camera(160), 1), ("realm-depth1", O.FLAGS_REALM, CAM.i_camera(160), 1))
    for name, flags, cam, depth in cases:
        for unit in (0, 5):
            st = assert_same(world, cam, 12, depth, 11, flags, unit=unit)
            assert st["exact_tests"] == st["segments"] * n and st["segments"] == st["samples"], (name, st)
            a = gpu(world, cam, 12, depth, seed=11, flags=flags, samples_per_unit=unit if unit else 12)
            b = gpu(world, cam, 12, depth, seed=11, flags=flags | _abi.F_LANE_KERNEL, samples_per_unit=unit if unit else 12)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2]["segments"] == b[2]["segments"], (name, unit)
        a = gpu(world, cam, 9, depth, seed=4, flags=flags, shard=(1, 3, 2))
        b = gpu(world, cam, 9, depth, seed=4, flags=flags | _abi.F_LANE_KERNEL, shard=(1, 3, 2))
        rows = render.shard_rows(cam.height, 1, 3, 2)
        assert np.array_equal(a[0][rows], b[0][rows]) and np.array_equal(a[1][rows], b[1][rows]) and a[2]["samples"] == b[2]["samples"], name


def test_cull_fuzz_against_exhaustive_scan():
    """The conservative fp32 cull + prefilter against the exhaustive fp64 scan on random scenes at
    scales from 1e-3 to 1e9, cameras inside, near and far, grazing rays, tiny and huge radii, clustered
    and coincident centres -- both kernels.  One missed sphere anywhere changes the image."""
    rng = np.random.default_rng(20261018)
    sp, mat = R.hittable.sphere, R.material
    for case in range(int(os.environ.get("RTCLJ_FUZZ_CASES", "48"))):   # a long soak: RTCLJ_FUZZ_CASES=1500
        scale = 10.0 ** rng.uniform(-3, 9)
        n = int(rng.choice([1, 2, 7, 31, 64, 200, 513]))
        spread = scale * 10.0 ** rng.uniform(-2, 2)
        offset = rng.normal(size=3) * scale * (0 if case % 3 else 1e3)   # a scene far from the origin
        world = []
        for i in range(n):
            c = offset + rng.normal(size=3) * spread * (0.01 if rng.random() < 0.2 else 1.0)
            if world and rng.random() < 0.1:
                c = np.array(world[-1]["hittable/center"])                  # coincident centres
            r = scale * 10.0 ** rng.uniform(-3, 1) * (-1 if rng.random() < 0.05 else 1)
            kind = rng.integers(0, 3)
            m = (mat.lambertian(tuple(rng.random(3))) if kind == 0 else
                 mat.metal(tuple(rng.random(3)), float(rng.random() * 1.2)) if kind == 1 else
                 mat.dielectric(float(rng.uniform(0.5, 2.5))))
            world.append(S.body(sp(tuple(float(x) for x in c), float(r)), m))
        pick = np.array(world[int(rng.integers(0, n))]["hittable/center"])
        where = rng.integers(0, 3)
        look_from = (pick + rng.normal(size=3) * scale * 1e-4 if where == 0 else        # inside / on a sphere
                     pick + rng.normal(size=3) * spread * 3 if where == 1 else           # among the spheres
                     offset + rng.normal(size=3) * spread * 10.0 ** rng.uniform(1, 4))   # far away
        cam = CAM.main_camera(48, 27, vfov=float(rng.uniform(1, 120)), look_from=tuple(float(x) for x in look_from),
                              look_at=tuple(float(x) for x in pick), defocus_angle=float(rng.choice([0.0, 2.0])),
                              focus_dist=float(np.linalg.norm(look_from - pick) + 1e-9 * scale))
        variant = O.FLAGS_REALM if case % 2 else O.FLAGS_MAIN   # forward / innermost-first product, Schlick on / off
        # every kernel takes its turn: the library's choice and the shared-memory-table kernel always, plus one of
        # the other small-scene kernels (two paths per lane, wavefront, dedicated cull warps)
        rotating = (_abi.F_LANE2_KERNEL, _abi.F_WAVE_KERNEL, _abi.F_SPLIT_KERNEL)[case % 3]
        for extra in (0, _abi.F_SMEM_TABLE) + ((rotating,) if n <= 512 else ()):
            a, ra, sa = gpu(world, cam, 4, 12, seed=case, flags=variant | extra, samples_per_unit=4)
            b, rb, sb = gpu(world, cam, 4, 12, seed=case, flags=variant | extra | _abi.F_NO_CULL, samples_per_unit=4)
            assert sa["segments"] == sb["segments"] and np.array_equal(a, b, equal_nan=True), (case, extra, n, scale)
        if case % 4 == 0:   # and the whole pipeline against the CPU oracle
            lo, ro, so = O.render(S.to_soa(world), cam, 4, 12, seed=case, flags=variant, threads=8, samples_per_unit=4)
            assert so.segments == sa["segments"] and np.array_equal(lo, a, equal_nan=True) and np.array_equal(ro, ra), (case, n, scale)


def test_both_kernels_agree_with_the_oracle():
    """Scenes of <= 512 spheres normally run the constant-table kernel; RTCLJ_F_SMEM_TABLE sends
    them through the shared-memory-table kernel (TMA staging, survivor lists, flushes) as well."""
    for world, cam, flags, seed in (
            (S.cover_hittables(7), CAM.main_camera(128, 72, **S.COVER_CAMERA), O.FLAGS_MAIN, 3),
            (S.realm_hittables(), CAM.realm_camera(96), O.FLAGS_REALM, 4),
            (S.cover_hittables(9)[:17], CAM.main_camera(64, 36, **S.COVER_CAMERA), O.FLAGS_MAIN, 5)):
        soa = S.to_soa(world)
        lin_o, rgb_o, st_o = O.render(soa, cam, 8, 50, seed=seed, flags=flags, threads=8, samples_per_unit=8)
        for extra in (0, _abi.F_SMEM_TABLE):
            lin_g, rgb_g, st_g = gpu(soa, cam, 8, 50, seed=seed, flags=flags | extra, samples_per_unit=8)
            assert np.array_equal(lin_o, lin_g) and np.array_equal(rgb_o, rgb_g), extra
            assert st_g["segments"] == st_o.segments


SMALL_KERNELS = {"lane": _abi.F_LANE_KERNEL, "lane2": _abi.F_LANE2_KERNEL, "wave": _abi.F_WAVE_KERNEL,
                 "split": _abi.F_SPLIT_KERNEL}


@pytest.mark.parametrize("which", sorted(SMALL_KERNELS))
def test_every_small_scene_kernel_agrees_with_the_oracle(which):
    """Scenes of <= 512 spheres have four kernels (one path per lane, two paths per lane, wavefront, dedicated cull warps);
    the library picks one, the RTCLJ_F_*_KERNEL flags select each.  All of them against the oracle:
    every variant, defocus, glass, depth caps (attenuation stack, K_END), chunked units, ragged sizes,
    exact ties, an empty list, shards, and the device-resident call."""
    extra = SMALL_KERNELS[which]
    sp, mat = R.hittable.sphere, R.material
    walls = [S.body(sp((0, 0, 0), 50.0), mat.metal((0.99, 0.98, 0.97), 0.0)),
             S.body(sp((0, 0, -1), 0.5), mat.dielectric(1.5)),
             S.body(sp((1.2, 0, -1), 0.5), mat.lambertian((0.5, 0.5, 0.9)))]
    tie = [S.body(sp((0, 0, -1), 0.5), mat.lambertian((0.9, 0.1, 0.1))),
           S.body(sp((0, 0, -1), 0.5), mat.lambertian((0.1, 0.9, 0.1)))]
    cases = [
        (S.cover_hittables(7), CAM.main_camera(160, 90, **S.COVER_CAMERA), 8, 50, 7, O.FLAGS_MAIN, 0),
        (S.cover_hittables(11), CAM.realm_camera(120, 67, look_from=(13.0, 2.0, 3.0), look_at=(0.0, 0.0, 0.0)), 9, 50, 22, O.FLAGS_REALM, 4),
        (S.main_hittables(), CAM.main_camera(200), 32, 50, 1, O.FLAGS_MAIN, 7),
        (S.realm_hittables(), CAM.realm_camera(200), 32, 50, 2, O.FLAGS_REALM, 0),
        (S.i_hittables(), CAM.i_camera(200), 16, 50, 3, O.FLAGS_I, 0),
        (S.realm_hittables(), CAM.i_camera(64), 4, 1, 2, O.FLAGS_REALM, 0),
        (S.cover_hittables(2)[:7], CAM.main_camera(37, 23, **S.COVER_CAMERA), 7, 50, 5, O.FLAGS_MAIN, 3),
        (tie, CAM.i_camera(48), 4, 50, 1, O.FLAGS_REALM, 0),
        ([], CAM.realm_camera(32), 4, 50, 1, O.FLAGS_REALM, 0),
        (S.cover_hittables(9)[:17], CAM.main_camera(1, 40, **S.COVER_CAMERA), 5, 50, 6, O.FLAGS_MAIN, 2),
    ]
    for depth in (1, 2, 7, 120):
        cases.append((walls, CAM.i_camera(40), 4, depth, 3, O.FLAGS_MAIN, 0))
        cases.append((walls, CAM.i_camera(40), 4, depth, 3, O.FLAGS_REALM, 0))
    for world, cam, spp, depth, seed, flags, unit in cases:
        st = assert_same(world, cam, spp, depth, seed, flags | extra, unit=unit)
    # shards: the union of the shards' rows is the whole image
    world, cam = S.main_hittables(), CAM.main_camera(96)
    whole, rgb_whole, st = gpu(world, cam, 8, 50, seed=5, samples_per_unit=4, flags=O.FLAGS_MAIN | extra)
    lin = np.zeros_like(whole); rgb = np.zeros_like(rgb_whole); segs = 0
    for idx in range(3):
        _, _, s1 = gpu(world, cam, 8, 50, seed=5, samples_per_unit=4, flags=O.FLAGS_MAIN | extra, shard=(idx, 3, 5),
                       out_linear=lin, out_rgb8=rgb)
        segs += s1["segments"]
    assert np.array_equal(lin, whole) and np.array_equal(rgb, rgb_whole) and segs == st["segments"]
    # cull == exhaustive fp64 scan
    a, ra, sa = gpu(S.cover_hittables(3), CAM.main_camera(96, 54, **S.COVER_CAMERA), 6, 50, seed=9, flags=O.FLAGS_MAIN | extra, samples_per_unit=6)
    b, rb, sb = gpu(S.cover_hittables(3), CAM.main_camera(96, 54, **S.COVER_CAMERA), 6, 50, seed=9, flags=O.FLAGS_MAIN | extra | _abi.F_NO_CULL, samples_per_unit=6)
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and sa["segments"] == sb["segments"]


def test_block_boundaries_and_kernel_switch():
    """Sphere counts around the cull-block sizes (16 / 32) and around 512, where the library
    switches from the constant-table kernel to the shared-memory-table kernel."""
    big = S.field_hittables(7, half=14)  # ~780 spheres, same generator as configs 3 and 5
    assert len(big) > 540
    cam = CAM.main_camera(48, 27, **S.COVER_CAMERA)
    for n in (15, 16, 17, 31, 32, 33, 47, 48, 49, 496, 511, 512, 513, 528, 529, 544):
        world = big[:n - 3] + big[-3:]  # keep the three big spheres in view
        assert len(world) == n
        assert_same(world, cam, 6, 50, n, O.FLAGS_MAIN, unit=4)


def test_cover_scene_medium_size_bit_exact():
    assert_same(S.cover_hittables(7), CAM.main_camera(320, 180, **S.COVER_CAMERA), 16, 50, 21, O.FLAGS_MAIN)
    assert_same(S.cover_hittables(11), CAM.realm_camera(240, 135, look_from=(13.0, 2.0, 3.0), look_at=(0.0, 0.0, 0.0)),
                16, 50, 22, O.FLAGS_REALM)


def test_far_origin_and_huge_spheres():
    """Numerics the survey flags (7.3-3): r = 1000 ground, rays leaving from far away."""
    world = S.cover_hittables(5)
    cam = CAM.main_camera(96, 54, vfov=40.0, look_from=(600.0, 30.0, 400.0), look_at=(0.0, 0.0, 0.0),
                          defocus_angle=0.0, focus_dist=10.0)
    assert_same(world, cam, 8, 50, 13, O.FLAGS_MAIN)
    # coordinates whose squares leave the fp32 range: such spheres are never culled, such origins scan everything
    sp, mat = R.hittable.sphere, R.material
    giant_ground = S.cover_hittables(5)[1:60] + [S.body(sp((0.0, -2e15 - 0.5, 0.0), 2e15), mat.lambertian((0.5, 0.6, 0.5))),
                                                 S.body(sp((4e16, 3e16, -9e16), 2e16), mat.metal((0.9, 0.9, 0.9), 0.0))]
    assert_same(giant_ground, CAM.main_camera(64, 36, **S.COVER_CAMERA), 8, 50, 14, O.FLAGS_MAIN)
    planet = [S.body(sp((0.0, 0.0, 0.0), 1e15), mat.lambertian((0.3, 0.5, 0.8))),
              S.body(sp((1.2e15, 0.0, 0.0), 1e14), mat.dielectric(1.5))]
    far_cam = CAM.main_camera(48, 27, vfov=60.0, look_from=(4e15, 1e15, 2e15), look_at=(0.0, 0.0, 0.0),
                              defocus_angle=0.0, focus_dist=10.0)
    assert_same(planet, far_cam, 8, 50, 15, O.FLAGS_MAIN)


def test_primary_ray_4k_bit_exact_8bit():
    """BASELINE.json config 4 at full 3840x2160: (i) the raytracing-i normal-shading scene,
    (ii) realm semantics at max-depth 1 on a camera that sees sky.  8-bit output bit-exact."""
    cam = CAM.i_camera(3840)
    assert (cam.width, cam.height) == (3840, 2160)
    for world, flags, depth, seed in ((S.i_hittables(), O.FLAGS_I, 50, 4), (S.realm_hittables(), O.FLAGS_REALM, 1, 6)):
        soa = S.to_soa(world)
        lin_o, rgb_o, st_o = O.render(soa, cam, 4, depth, seed=seed, flags=flags, threads=8)
        lin_g, rgb_g, st_g = gpu(soa, cam, 4, depth, seed=seed, flags=flags, samples_per_unit=4)
        assert np.array_equal(rgb_o, rgb_g)
        assert np.array_equal(lin_o, lin_g)
        assert st_g["segments"] == st_o.segments == 3840 * 2160 * 4


def test_edge_cases():
    cam = CAM.realm_camera(32)
    # empty hittable list: sky only
    assert_same([], cam, 4, 50, 1, O.FLAGS_REALM)
    # one sphere, one pixel, one sample
    one = [S.body(R.hittable.sphere((0, 0, -1), 0.5), R.material.metal((0.9, 0.9, 0.9), 0.0))]
    assert_same(one, CAM.i_camera(1, 1), 1, 50, 1, O.FLAGS_REALM)
    # depth 0 -> black, nothing traced; depth 1 -> sky where the primary ray misses
    lin, rgb, st = gpu(S.realm_hittables(), cam, 4, 0, flags=O.FLAGS_REALM)
    assert not lin.any() and not rgb.any() and st["segments"] == 0
    assert_same(S.realm_hittables(), CAM.i_camera(64), 4, 1, 2, O.FLAGS_REALM)
    # ragged sizes: n not a multiple of 4, odd image sizes, spp not a multiple of the unit
    for n in (1, 2, 3, 5, 6, 7):
        world = S.cover_hittables(2)[:n]
        assert_same(world, CAM.main_camera(37, 23, **S.COVER_CAMERA), 7, 50, n, O.FLAGS_MAIN, unit=3)
    # concentric spheres and exact ties: the first body wins
    tie = [S.body(R.hittable.sphere((0, 0, -1), 0.5), R.material.lambertian((0.9, 0.1, 0.1))),
           S.body(R.hittable.sphere((0, 0, -1), 0.5), R.material.lambertian((0.1, 0.9, 0.1)))]
    assert_same(tie, CAM.i_camera(48), 4, 50, 1, O.FLAGS_REALM)


def test_parameters_outside_the_reference_scenes():
    """Values the reference's constructors accept without checks (hittable.clj:7, material.clj:13-36):
    negative radius (the book's hollow-glass trick: flips the outward normal), zero radius, fuzz > 1,
    ior = 1 and ior < 1, albedo outside [0,1], coincident centres with different radii."""
    sp, mat = R.hittable.sphere, R.material
    world = [S.body(sp((0, -100.5, -1), 100.0), mat.lambertian((0.8, 0.8, 0.0))),
             S.body(sp((-1.0, 0, -1), 0.5), mat.dielectric(1.5)),
             S.body(sp((-1.0, 0, -1), -0.4), mat.dielectric(1.5)),          # hollow glass, book style
             S.body(sp((0.0, 0, -1.2), 0.5), mat.metal((1.3, 0.6, -0.1), 1.7)),
             S.body(sp((1.0, 0, -1), 0.5), mat.dielectric(1.0)),
             S.body(sp((1.0, 0.9, -1), 0.3), mat.dielectric(0.4)),
             S.body(sp((0.3, 0.1, -0.4), 0.0), mat.lambertian((0.5, 0.5, 0.5))),
             S.body(sp((0.0, 0.8, -1.2), -0.25), mat.lambertian((0.2, 0.9, 0.4))),
             S.body(sp((0.0, 0.8, -1.2), 0.1), mat.metal((0.9, 0.9, 0.9), 0.0))]
    for flags in (O.FLAGS_MAIN, O.FLAGS_REALM):
        assert_same(world, CAM.main_camera(80), 16, 50, 21, flags)
        assert_same(world, CAM.i_camera(64), 8, 50, 22, flags)
    # camera inside a negative-radius sphere
    inside = [S.body(sp((0, 0, 0), -30.0), mat.lambertian((0.7, 0.7, 0.7))),
              S.body(sp((0, 0, -1), 0.5), mat.dielectric(1.5))]
    assert_same(inside, CAM.i_camera(48), 8, 50, 23, O.FLAGS_MAIN)


def test_depth_cap_and_deep_paths():
    # a mirror box: paths bounce until the depth cap; exercises the reverse-product stack
    walls = [S.body(R.hittable.sphere((0, 0, 0), 50.0), R.material.metal((0.99, 0.98, 0.97), 0.0)),
             S.body(R.hittable.sphere((0, 0, -1), 0.5), R.material.dielectric(1.5)),
             S.body(R.hittable.sphere((1.2, 0, -1), 0.5), R.material.lambertian((0.5, 0.5, 0.9)))]
    for depth in (1, 2, 7, 50, 120):
        assert_same(walls, CAM.i_camera(40), 4, depth, 3, O.FLAGS_MAIN)
        assert_same(walls, CAM.i_camera(40), 4, depth, 3, O.FLAGS_REALM)


def test_sharded_rows_union_equals_whole():
    world, cam = S.main_hittables(), CAM.main_camera(96)
    whole, rgb_whole, st = gpu(world, cam, 8, 50, seed=5, samples_per_unit=8)
    for count, rows in ((2, 4), (3, 5), (8, 1), (4, 64)):
        lin = np.zeros_like(whole)
        rgb = np.zeros_like(rgb_whole)
        segs = 0
        for idx in range(count):
            _, _, s = gpu(world, cam, 8, 50, seed=5, samples_per_unit=8, shard=(idx, count, rows),
                          out_linear=lin, out_rgb8=rgb)
            segs += s["segments"]
        assert np.array_equal(lin, whole) and np.array_equal(rgb, rgb_whole) and segs == st["segments"]


def test_chunked_units_match_oracle_and_reference_order():
    world, cam = S.main_hittables(), CAM.main_camera(64)
    assert_same(world, cam, 30, 50, 8, O.FLAGS_MAIN, unit=7)
    strict, _, _ = gpu(world, cam, 30, 50, seed=8, samples_per_unit=30)
    auto, rgb_auto, st = gpu(world, cam, 30, 50, seed=8)  # library-chosen units
    assert st["samples_per_unit"] >= 1
    assert np.allclose(auto, strict, rtol=1e-13, atol=1e-300)


def test_full_size_properties_config2():
    """BASELINE.json config 2 (1920x1080, 100 spp, depth 50) is too big for the oracle; check
    size-independent properties: determinism, seed sensitivity, quantisation consistency,
    a band of rows against the oracle, and the segment statistics."""
    world, cam = S.main_hittables(), CAM.main_camera(1920)
    assert cam.height == 1080
    lin, rgb, st = gpu(world, cam, 100, 50, seed=1)
    lin2, rgb2, st2 = gpu(world, cam, 100, 50, seed=1)
    assert np.array_equal(lin, lin2) and st["segments"] == st2["segments"]
    assert np.array_equal(render.quantise_rgb8(lin), rgb)
    assert abs(st["segments"] / st["samples"] - 3.675) < 0.01
    assert st["samples"] == 1920 * 1080 * 100
    rows = (536, 540)
    lin_o, rgb_o, _ = O.render(S.to_soa(world), cam, 100, 50, seed=1, flags=O.FLAGS_MAIN, threads=8, rows=rows,
                               samples_per_unit=st["samples_per_unit"])
    assert np.array_equal(lin_o[rows[0]:rows[1]], lin[rows[0]:rows[1]])
    assert np.array_equal(rgb_o[rows[0]:rows[1]], rgb[rows[0]:rows[1]])


def test_extreme_image_shapes():
    """Maximum sizes the reference could be asked for: a 7680x4320 frame (33 M pixels), one-pixel-wide
    and one-pixel-high images, and a small image at 4000 spp (chunked sums).  Rows / whole images
    against the oracle."""
    world = S.main_hittables()
    cam = CAM.main_camera(7680, 4320)
    lin, rgb, st = gpu(world, cam, 1, 6, seed=3)
    assert st["samples"] == 7680 * 4320 and st["samples_per_unit"] == 1
    for j in (0, 2161, 4319):
        lin_o, rgb_o, _ = O.render(S.to_soa(world), cam, 1, 6, seed=3, flags=O.FLAGS_MAIN, threads=8, rows=(j, j + 1))
        assert np.array_equal(lin_o[j], lin[j]) and np.array_equal(rgb_o[j], rgb[j])
    assert render.encode_ppm(rgb, device=0) == render.encode_ppm(rgb)      # 33 M lines of text, both writers
    del lin, rgb
    assert_same(world, CAM.main_camera(1, 300), 6, 50, 4, O.FLAGS_MAIN)
    assert_same(world, CAM.main_camera(700, 1), 6, 50, 5, O.FLAGS_MAIN)
    soa = S.to_soa(world)
    small = CAM.main_camera(12, 7)
    lin_g, rgb_g, st_g = gpu(soa, small, 4000, 50, seed=6)
    lin_o, rgb_o, st_o = O.render(soa, small, 4000, 50, seed=6, flags=O.FLAGS_MAIN, threads=8,
                                  samples_per_unit=st_g["samples_per_unit"])
    assert np.array_equal(lin_o, lin_g) and np.array_equal(rgb_o, rgb_g) and st_g["segments"] == st_o.segments


def test_device_resident_context_matches_host_call():
    import torch
    world, cam = S.cover_hittables(7), CAM.main_camera(128, 72, **S.COVER_CAMERA)
    ref, rgb_ref, st_ref = gpu(world, cam, 8, 50, seed=4, samples_per_unit=8)
    ctx = render.Context(0)
    ctx.set_scene(world)
    out = torch.zeros((72, 128, 3), dtype=torch.float64, device="cuda:0")
    out8 = torch.zeros((72, 128, 3), dtype=torch.uint8, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    ctx.render(cam, 8, 50, seed=4, samples_per_unit=8, d_out_linear=out.data_ptr(), d_out_rgb8=out8.data_ptr(),
               stream=stream)
    st = ctx.stats(stream)
    assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(out8.cpu().numpy(), rgb_ref)
    assert st["segments"] == st_ref["segments"]
    ctx.close()


def test_concurrent_host_calls_from_several_threads():
    """JVM hosts call from pool threads (raytracing.clj:162-166): the host-buffer entry points share
    one cached context per device and must serialise themselves."""
    import threading
    world, cam = S.main_hittables(), CAM.main_camera(64)
    want = {seed: gpu(world, cam, 4, 50, seed=seed) for seed in (1, 2, 3, 4)}
    got, errors = {}, []

    def work(seed):
        try:
            for _ in range(3):
                if seed % 2:
                    got[seed] = gpu(world, cam, 4, 50, seed=seed)
                else:  # the one-shard-per-call entry point, two shards
                    lin = np.zeros((cam.height, cam.width, 3)); rgb = np.zeros((cam.height, cam.width, 3), dtype=np.uint8)
                    for idx in range(2):
                        gpu(world, cam, 4, 50, seed=seed, shard=(idx, 2, 3), out_linear=lin, out_rgb8=rgb)
                    got[seed] = (lin, rgb, None)
                render.encode_ppm(want[seed][1], device=0)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(s,)) for s in want]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for seed in want:
        assert np.array_equal(got[seed][0], want[seed][0]) and np.array_equal(got[seed][1], want[seed][1])


def test_plain_c_client_renders_the_same_image(tmp_path):
    """A C program (not Python) driving rtclj_render: the reference's scene, compared with the oracle."""
    import subprocess
    from test_abi import build_c_client
    dump = str(tmp_path / "linear.f64")
    out = subprocess.run([build_c_client(tmp_path), "gpu", dump], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "ok gpu"
    cam = CAM.main_camera(64, 36)
    lin_o, _, _ = O.render(S.to_soa(S.main_hittables()), cam, 8, 50, seed=1, flags=O.FLAGS_MAIN, threads=4, samples_per_unit=8)
    got = np.fromfile(dump, dtype=np.float64).reshape(36, 64, 3)
    assert np.array_equal(got, lin_o)


def test_entry_points_write_the_files_the_reference_writes(tmp_path, monkeypatch):
    """`clojure -M:main 8 50` / `-M:realm` / raytracing-i: scene.ppm + scene.png, scene-realm.ppm, scene-i.ppm."""
    from PIL import Image
    from raytracing_clj_b200 import main as entry
    monkeypatch.chdir(tmp_path)
    st = entry.main_variant(8, 50)
    ppm = render.decode_ppm(open("scene.ppm", "rb").read())
    assert ppm.shape == (225, 400, 3) and st["samples"] == 400 * 225 * 8
    assert np.array_equal(np.asarray(Image.open("scene.png").convert("RGB")), ppm)
    _, rgb_o, _ = O.render(S.to_soa(S.main_hittables()), CAM.main_camera(), 8, 50, seed=1, flags=O.FLAGS_MAIN, threads=8,
                           samples_per_unit=st["samples_per_unit"])
    assert np.array_equal(ppm, rgb_o)
    entry.i_variant()
    assert render.decode_ppm(open("scene-i.ppm", "rb").read()).shape == (224, 400, 3)


def test_non_finite_inputs_are_refused():
    world = S.main_hittables()
    for field, value in (("hittable/center", (float("nan"), 0.0, -1.0)), ("hittable/radius", float("inf")),
                         ("material/albedo", (0.5, float("nan"), 0.5))):
        bad = [dict(b) for b in world]
        bad[1][field] = value
        with pytest.raises(_abi.RtcljError) as e:
            gpu(bad, CAM.main_camera(16), 1, 5)
        assert e.value.code == _abi.E_INVALID
    import dataclasses
    cam = dataclasses.replace(CAM.main_camera(16), center=(float("inf"), 0.0, 0.0))
    with pytest.raises(_abi.RtcljError) as e:
        gpu(world, cam, 1, 5)
    assert e.value.code == _abi.E_INVALID


def test_error_behaviour():
    import ctypes as C
    cam = CAM.main_camera(16)
    with pytest.raises(_abi.RtcljError) as e:
        gpu(S.main_hittables(), cam, 0, 5)
    assert e.value.code == _abi.E_INVALID
    bad = S.to_soa(S.main_hittables())
    bad[2][0] = 9
    with pytest.raises(_abi.RtcljError):
        gpu(bad, cam, 1, 5)
    with pytest.raises(_abi.RtcljError) as e:
        gpu(S.main_hittables(), cam, 1, 5, devices=[99])
    assert e.value.code == _abi.E_INVALID
    big = S._random_field(1, -130, 130)  # > 65 532 spheres: survivor entries hold 16-bit block indices
    assert len(big) > 65532
    with pytest.raises(_abi.RtcljError) as e:
        gpu(big, cam, 1, 5)
    assert e.value.code == _abi.E_TOO_LARGE


def test_scene_larger_than_shared_memory():
    """~25 600 spheres: 410 KB of cull table, of which ~190 KB sit in shared memory and the rest is
    read from global memory."""
    world = S._random_field(1, -80, 80)
    assert len(world) > 25000
    cam = CAM.main_camera(40, 22, vfov=40.0, look_from=(95.0, 20.0, 30.0), look_at=(0.0, 0.0, 0.0),
                          defocus_angle=0.3, focus_dist=90.0)
    assert_same(world, cam, 2, 50, 17, O.FLAGS_MAIN)
