"""Pins the CPU oracle (oracle/rt_oracle.c) before anything trusts it.

The reference has no tests and no seeded output; its only result pins
are two committed renders.  They are STATISTICAL goldens (unseeded RNG), so the gates
are the ones BASELINE.md states: channel means within 0.1/255 of the golden, and
MAE against the golden within 5 % of the oracle's own seed-to-seed MAE."""
import os

import numpy as np
import pytest

import oracle_lib as O
import raytracing_clj_b200 as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ref_images():
    return np.load(os.path.join(GOLD, "reference_images.npz"))


def test_philox_known_answers():
    # SURVEY.md Appendix E (Random123 known-answer vectors)
    assert O.philox((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert O.philox((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert O.philox((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_uniform_is_24_bit_and_in_range():
    us = [O.lib().rto_uniform(9, p, s, 0, 0, w) for p in range(20) for s in range(5) for w in range(4)]
    assert all(0.0 <= u < 1.0 for u in us)
    assert all(float(u * 2 ** 24).is_integer() for u in us)
    assert len(set(us)) > 390


def test_image_dimensions_match_reference_headers(ref_images):
    # 400x225 (main, exact Ratio arithmetic) vs 400x224 (realm, Ratio->double quirk)
    assert ref_images["scene_main"].shape == (225, 400, 3)
    assert ref_images["scene_realm"].shape == (224, 400, 3)
    assert (R.camera.main_camera().width, R.camera.main_camera().height) == (400, 225)
    assert (R.camera.realm_camera().width, R.camera.realm_camera().height) == (400, 224)
    assert R.camera.i_camera().height == 224
    # the quirk only bites at W=400
    assert [R.camera.image_height_realm(w) for w in (1200, 1920, 3840)] == [675, 1080, 2160]


def _stat_gate(rgb_a, rgb_b, gold):
    mean_a = rgb_a.reshape(-1, 3).mean(0)
    mean_g = gold.reshape(-1, 3).astype(np.float64).mean(0)
    assert np.all(np.abs(mean_a - mean_g) < 0.1), (mean_a, mean_g)
    mae_gold = np.abs(rgb_a.astype(int) - gold.astype(int)).mean()
    mae_seed = np.abs(rgb_a.astype(int) - rgb_b.astype(int)).mean()
    assert abs(mae_gold - mae_seed) < 0.05 * mae_seed, (mae_gold, mae_seed)


def test_main_variant_reproduces_scene_ppm(ref_images):
    soa = R.scenes.to_soa(R.scenes.main_hittables())
    cam = R.camera.main_camera()
    _, a, st = O.render(soa, cam, 100, 50, seed=1, flags=O.FLAGS_MAIN, threads=8)
    _, b, _ = O.render(soa, cam, 100, 50, seed=2, flags=O.FLAGS_MAIN, threads=8)
    _stat_gate(a, b, ref_images["scene_main"])
    # workload statistics measured by the survey
    assert abs(st.segments / st.samples - 3.675) < 0.01
    assert st.samples == 400 * 225 * 100


def test_realm_variant_reproduces_scene_realm_ppm(ref_images):
    soa = R.scenes.to_soa(R.scenes.realm_hittables())
    cam = R.camera.realm_camera()
    _, a, st = O.render(soa, cam, 100, 50, seed=1, flags=O.FLAGS_REALM, threads=8)
    _, b, _ = O.render(soa, cam, 100, 50, seed=2, flags=O.FLAGS_REALM, threads=8)
    _stat_gate(a, b, ref_images["scene_realm"])
    assert abs(st.segments / st.samples - 3.714) < 0.01


def test_repl_known_answers():
    # the reference's only worked examples (REPL scratch): realm/vec3.clj:159-164 adds
    # (0,1,1)+(0,1,1) and takes lengthSquared -> 8.0; vec3i.clj:81-87 normalises (3,-2,0.5).
    # Exercised through hit_anything: a unit sphere at the origin hit from (3,-2,0.5)*2
    # along -(3,-2,0.5) returns the outward normal = unit((3,-2,0.5)).
    soa = R.scenes.to_soa([R.scenes.body(R.hittable.sphere((0, 0, 0), 1.0), R.material.lambertian((1, 1, 1)))])
    v = np.array([3.0, -2.0, 0.5])
    idx, t, p, n, ff = O.hit_anything(soa, tuple(2 * v), tuple(-v))
    L = np.sqrt(3.0 * 3.0 + -2.0 * -2.0 + 0.5 * 0.5)
    assert idx == 0 and ff
    np.testing.assert_allclose(n, v / L, rtol=0, atol=1e-15)
    w = np.array([0.0, 1.0, 1.0]) + np.array([0.0, 1.0, 1.0])
    assert float(w @ w) == 8.0


def test_hit_anything_semantics():
    S, M = R.scenes, R.material
    two = S.to_soa([S.body(R.hittable.sphere((0, 0, -1), 0.5), M.lambertian((1, 1, 1))),
                    S.body(R.hittable.sphere((0, 0, -1), 0.5), M.metal((1, 1, 1), 0.0))])
    # exact tie: the FIRST body wins (strict bounds, raytracing.clj:33-43)
    assert O.hit_anything(two, (0, 0, 0), (0, 0, -1))[0] == 0
    # t is in units of the UN-normalised direction (Appendix B.2)
    assert O.hit_anything(two, (0, 0, 0), (0, 0, -2))[1] == 0.25
    # origin on the surface: near root <= t_min, so the far root is taken
    idx, t, p, n, ff = O.hit_anything(two, (0, 0, -0.5), (0, 0, -1))
    assert idx == 0 and t == 1.0 and not ff and n == (0.0, 0.0, 1.0)
    # miss
    assert O.hit_anything(two, (0, 0, 0), (0, 1, 0))[0] == -1
    # empty list
    empty = S.to_soa([])
    assert O.hit_anything(empty, (0, 0, 0), (0, 0, -1))[0] == -1


def test_quantise_edges():
    q = O.lib().rto_quantise
    assert q(0.0, 0) == 0 and q(-1.0, 0) == 0 and q(float("nan"), 0) == 0
    assert q(1.0, 0) == 255 and q(50.0, 0) == 255  # clamp 0.999 -> int(255.744)
    assert q(0.25, 0) == 128
    assert q(1.0, 1) == 255 and q(0.5, 1) == 127 and q(float("nan"), 1) == 0


def test_depth_semantics():
    soa = R.scenes.to_soa(R.scenes.realm_hittables())
    # default camera: every primary ray hits something, so depth 1 is black (Appendix A.3)
    lin, _, st = O.render(soa, R.camera.realm_camera(40), 4, 1, flags=O.FLAGS_REALM)
    assert st.segments == st.samples and not lin.any()
    # depth 0 traces nothing
    lin, _, st = O.render(soa, R.camera.realm_camera(40), 4, 0, flags=O.FLAGS_REALM)
    assert st.segments == 0 and not lin.any()
    # a camera that sees sky: depth 1 = sky where the primary ray misses
    lin, _, _ = O.render(soa, R.camera.i_camera(40), 4, 1, flags=O.FLAGS_REALM)
    assert lin[0].min() > 0.4 and lin.max() <= 1.0


def test_threads_rows_and_chunking_invariants():
    soa = R.scenes.to_soa(R.scenes.main_hittables())
    cam = R.camera.main_camera(48)
    a, ra, _ = O.render(soa, cam, 12, 50, seed=4, flags=O.FLAGS_MAIN, threads=1)
    b, rb, _ = O.render(soa, cam, 12, 50, seed=4, flags=O.FLAGS_MAIN, threads=5)
    assert np.array_equal(a, b) and np.array_equal(ra, rb)  # counter-based RNG: partition-independent
    c, _, _ = O.render(soa, cam, 12, 50, seed=4, flags=O.FLAGS_MAIN, rows=(5, 9))
    assert np.array_equal(c[5:9], a[5:9]) and not c[:5].any() and not c[9:].any()
    d, _, _ = O.render(soa, cam, 12, 50, seed=4, flags=O.FLAGS_MAIN, samples_per_unit=5)
    assert np.allclose(d, a, rtol=1e-13, atol=0) and np.abs(d - a).max() < 1e-15 * 12


def test_committed_oracle_fixtures_still_reproduce():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    fx = np.load(os.path.join(GOLD, "oracle_fixtures.npz"))
    for name, (bodies, cam, spp, depth, seed, flags, unit) in mg.fixture_cases().items():
        lin, rgb, st = O.render(R.scenes.to_soa(bodies), cam, spp, depth, seed=seed, flags=flags,
                                samples_per_unit=unit)
        assert np.array_equal(lin, fx[name + "/linear"]), name
        assert np.array_equal(rgb, fx[name + "/rgb8"]), name
        assert st.segments == int(fx[name + "/segments"][0]), name
