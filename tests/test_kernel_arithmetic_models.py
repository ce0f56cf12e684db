"""CPU models of two pieces of device arithmetic whose correctness the kernels rely on (no GPU needed): the
packed prefilter candidate (rtclj_kernels.cuh: cand_key / cand_insert) and the operand window of the
shared-reciprocal division (div_by / divs_by).  The models restate the device code operation for operation in
numpy float32 / integer arithmetic; the GPU side of both is covered by the parity tests and by
tools/microbench/div_recip_check.cu (profiles/r2_div_recip_check.txt)."""
import numpy as np

F32 = np.float32


def cand_key(lo, i):
    """rtclj_kernels.cuh cand_key(): clamp, lower by 2^-13 of itself (one fused multiply-add), index into the nine
    low mantissa bits."""
    lk = np.maximum(lo.astype(F32), F32(-1.0e38))
    # fmaf(-|lk|, 2^-13, lk): the product and the sum are exact in double, so one rounding to float32 = the fused result
    lk = (lk.astype(np.float64) - np.abs(lk.astype(np.float64)) * 2.0 ** -13).astype(F32)
    lk = lk - F32(1.0e-30)   # covers |lo| < 2^-113, where the relative step underflows
    bits = (lk.view(np.uint32) & np.uint32(0xFFFFFE00)) | i.astype(np.uint32)
    return bits.view(F32)


def test_packed_candidate_is_a_lower_bound_that_carries_its_index():
    rng = np.random.default_rng(7)
    n = 400_000
    mag = 10.0 ** rng.uniform(-30, 30, n)
    lo = (rng.choice([-1.0, 1.0], n) * mag).astype(F32)
    lo[:1000] = F32(0.0)
    lo[1000:2000] = F32(-np.inf)           # an fp32 overflow in the bound
    lo[2000:2100] = np.finfo(F32).tiny     # smallest normal
    idx = rng.integers(0, 512, n)
    key = cand_key(lo, idx)
    assert np.all(np.isfinite(key))
    assert np.array_equal(key.view(np.uint32) & np.uint32(0x1FF), idx.astype(np.uint32))
    # still a lower bound, whatever the index bits (-inf, "no information", becomes -1e38: below every root, which
    # the input range keeps under 1e30 in magnitude -- DESIGN.md 4.8)
    assert np.all(key <= np.maximum(lo, F32(-1.0e38)))
    finite = np.isfinite(lo) & (np.abs(lo) > 1e-20)
    assert np.all(np.abs(key[finite] - lo[finite]) <= np.abs(lo[finite]) * F32(2.0 ** -12))   # and a tight one
    assert np.all(key < F32(1.0e38))                          # never mistaken for an empty slot (3.0e38)


def test_candidate_insertion_keeps_the_three_smallest_sorted():
    rng = np.random.default_rng(11)
    for _ in range(300):
        m = int(rng.integers(1, 9))
        lo = rng.normal(0, 50, m).astype(F32)
        idx = rng.permutation(512)[:m]
        keys = cand_key(lo, idx)
        k = [F32(3.0e38)] * 3
        displaced = []
        for key in keys:                                      # cand_insert(): a float min / max pair per slot
            for s in range(3):
                t = min(k[s], key); key = max(k[s], key); k[s] = t
            if key < F32(1.0e38):
                displaced.append(key)
        want = sorted(keys.tolist())
        assert [float(x) for x in k[: min(m, 3)]] == want[: min(m, 3)]
        assert sorted(float(x) for x in displaced) == want[3:]
        for s in range(min(m, 3)):                            # the slot still names its sphere
            assert int(np.array([k[s]], dtype=F32).view(np.uint32)[0] & 0x1FF) in set(int(x) for x in idx)


def exp_off(v):
    """rtclj_kernels.cuh exp_off(): biased exponent of a double, offset so that ONE unsigned compare tests a window."""
    hi = (np.atleast_1d(np.asarray(v, dtype=np.float64)).view(np.uint64) >> np.uint64(32)).astype(np.uint64)
    off = ((hi & np.uint64(0x7FF00000)) + np.uint64(2 ** 32 - (543 << 20))) & np.uint64(0xFFFFFFFF)   # mod 2^32, as on the device
    return off.astype(np.uint32) if np.ndim(v) else np.uint32(off[0])


K_EXP_SPAN = np.uint32(961 << 20)


def test_division_window_on_the_operands_bounds_the_quotient():
    """Both operands inside [2^-480, 2^481) => the quotient inside (2^-961, 2^961), i.e. inside the range on which
    the fast path was verified against `/` -- so the kernels do not test the quotient."""
    rng = np.random.default_rng(3)
    n = 300_000
    e = rng.integers(-1074, 1024, (2, n))
    m = rng.uniform(1.0, 2.0, (2, n))
    with np.errstate(over="ignore", under="ignore"):
        a = np.ldexp(m[0], e[0]) * rng.choice([-1.0, 1.0], n)
        b = np.ldexp(m[1], e[1]) * rng.choice([-1.0, 1.0], n)
    a[:50] = 0.0; a[50:100] = -0.0; a[100:150] = np.inf; a[150:200] = np.nan
    with np.errstate(all="ignore"):
        ok = (exp_off(a) < K_EXP_SPAN) & (exp_off(b) < K_EXP_SPAN)
        q = a / b
    assert ok.sum() > 10_000
    assert np.all(np.isfinite(q[ok]))
    assert np.all((np.abs(q[ok]) > 2.0 ** -961) & (np.abs(q[ok]) < 2.0 ** 961))
    lo_edge, hi_edge = 2.0 ** -480, np.nextafter(2.0 ** 481, 0)
    assert exp_off(lo_edge) < K_EXP_SPAN and exp_off(hi_edge) < K_EXP_SPAN
    assert not exp_off(np.nextafter(lo_edge, 0)) < K_EXP_SPAN and not exp_off(2.0 ** 481) < K_EXP_SPAN
    for special in (0.0, -0.0, 5e-324, np.inf, -np.inf, np.nan):   # zero, denormal, inf, NaN: outside, they divide
        assert not exp_off(special) < K_EXP_SPAN


def test_zero_numerator_on_the_fast_path_has_the_quotients_sign():
    """div_by(): a zero numerator returns n * y with y ~ 1/d, sign(y) = sign(d): the exact signed zero of n / d."""
    for n in (0.0, -0.0):
        for d in (3.0, -3.0, 1e-300, -1e300):
            y = 1.0 / d                                        # any value with the sign of d
            assert np.signbit(n * y) == np.signbit(np.float64(n) / np.float64(d))
