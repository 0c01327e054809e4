"""The arithmetic argument behind `rsb_gemm_planes`, checked on the CPU: an fp32 number splits exactly enough into
three bf16 terms, and the six term products >= 2^-16 (the three "bands" the kernel issues as tensor-core MMAs)
reproduce the fp32 product to below fp32 rounding - the three dropped products are <= 2^-24 relative.
(The kernel's own accuracy is measured on the GPU in tests/test_gpu_gemm.py against fp64.)"""
import numpy as np


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32, by bit manipulation."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split3(x: np.ndarray):
    a0 = bf16_round(x)
    a1 = bf16_round((x - a0).astype(np.float32))
    a2 = bf16_round((x - a0 - a1).astype(np.float32))
    return a0, a1, a2


def test_three_bf16_terms_carry_an_fp32_mantissa():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    a0, a1, a2 = split3(x)
    rec = a0.astype(np.float64) + a1.astype(np.float64) + a2.astype(np.float64)
    rel = np.abs(rec - x.astype(np.float64)) / np.abs(x.astype(np.float64))
    assert rel.max() <= 2.0 ** -24                                  # 3 x 8 mantissa bits
    assert np.all(np.abs(a1) <= np.abs(a0) * 2.0 ** -8 * 1.01) and np.all(np.abs(a2) <= np.abs(a0) * 2.0 ** -16 * 1.01)


def test_six_term_products_match_the_fp32_gemm_to_below_fp32_rounding():
    rng = np.random.default_rng(1)
    m, n, k = 64, 48, 400
    a = rng.standard_normal((m, k)).astype(np.float32)
    b = rng.standard_normal((k, n)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    sa, sb = split3(a), split3(b)
    bands = {3: [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)],    # what the kernel computes (6 MMAs)
             5: [(i, j) for i in range(3) for j in range(3)]}        # all 9
    err = {}
    for nb, pairs in bands.items():
        acc = np.zeros((m, n))
        for i, j in pairs:
            acc += sa[i].astype(np.float64) @ sb[j].astype(np.float64)   # products exact, accumulation in fp64:
        err[nb] = np.linalg.norm(acc - ref) / np.linalg.norm(ref)        # isolates the error of the split itself
    fp32 = np.linalg.norm((a @ b).astype(np.float64) - ref) / np.linalg.norm(ref)   # a plain fp32 GEMM, for scale
    assert err[5] < 2.0 ** -24 and err[3] < 2.0 ** -22
    assert err[3] < fp32                                             # dropping 3 bands costs less than fp32 rounding
    # the dropped terms alone: a1*b2 + a2*b1 + a2*b2
    dropped = sum(sa[i].astype(np.float64) @ sb[j].astype(np.float64) for i, j in [(1, 2), (2, 1), (2, 2)])
    assert np.linalg.norm(dropped) / np.linalg.norm(ref) < 2.0 ** -22


# ------------------------------------------------------------------ RSB_PLANES_FP16X2 ---
def plane_scale(amax: float, max_exp: int = 40) -> float:
    """csrc/common.cuh::plane_scale: the power of two that puts amax into [2^13, 2^14), at most 2^max_exp."""
    if not (amax > 0.0) or not np.isfinite(amax):
        return 1.0
    _, e = np.frexp(np.float32(amax))                 # amax in [2^(e-1), 2^e)
    return float(np.ldexp(1.0, min(14 - int(e), max_exp)))


def split2_fp16(x: np.ndarray, s: float):
    xs = (x.astype(np.float32) * np.float32(s)).astype(np.float32)
    h0 = xs.astype(np.float16)
    h1 = (xs - h0.astype(np.float32)).astype(np.float16)
    return h0, h1


def test_plane_scale_puts_the_bound_below_fp16_overflow_and_is_a_power_of_two():
    for amax in [1.0, 0.999, 1.001, 3e-7, 65504.0, 1e30, 2.0 ** -60, 7.3e-3, 16384.0, 16383.9]:
        for cap in (14, 40):
            s = plane_scale(amax, cap)
            assert s == 2.0 ** round(np.log2(s))
            assert amax * s < 2.0 ** 14                           # fp16 max is 65504: two more bits of headroom
            if np.log2(s) < cap:
                assert amax * s >= 2.0 ** 13
    assert plane_scale(0.0) == 1.0 and plane_scale(float("nan")) == 1.0 and plane_scale(float("inf")) == 1.0


def test_two_scaled_fp16_planes_keep_22_bits_of_the_large_elements_and_an_absolute_floor_for_all():
    rng = np.random.default_rng(2)
    for mag in (1.0, 3e-7, 2e4):
        x = (rng.standard_normal(100000) * mag * np.exp(rng.uniform(-12, 0, 100000))).astype(np.float32)
        s = plane_scale(float(np.abs(x).max()))
        h0, h1 = split2_fp16(x, s)
        rec = (h0.astype(np.float64) + h1.astype(np.float64)) / s
        err = np.abs(rec - x.astype(np.float64))
        amax = float(np.abs(x).max())
        # every element: relative 2^-22 or the fp16 subnormal step / scale, whichever is larger
        assert np.all(err <= np.maximum(np.abs(x) * 2.0 ** -22, 2.0 ** -25 / s * 1.0001))
        assert err.max() <= amax * 2.0 ** -22


def test_three_fp16_plane_products_match_fp64_to_fp32_gemm_accuracy():
    rng = np.random.default_rng(3)
    m, n, k = 64, 48, 624
    for mag_a, mag_b in ((1.0, 0.05), (3e-7, 0.05), (50.0, 1e-3)):
        a = (rng.standard_normal((m, k)) * mag_a).astype(np.float32)
        b = (rng.standard_normal((k, n)) * mag_b).astype(np.float32)
        ref = a.astype(np.float64) @ b.astype(np.float64)
        sa, sb = plane_scale(float(np.abs(a).max())), plane_scale(float(np.abs(b).max()))
        a0, a1 = (h.astype(np.float64) for h in split2_fp16(a, sa))
        b0, b1 = (h.astype(np.float64) for h in split2_fp16(b, sb))
        acc = (a0 @ b1 + a1 @ b0 + a0 @ b0) / (sa * sb)               # the kernel's three MMAs, smallest first
        err = np.abs(acc - ref).max() / np.abs(ref).max()
        fp32 = np.abs((a @ b).astype(np.float64) - ref).max() / np.abs(ref).max()
        assert err < 2.5e-7 and err < fp32, (err, fp32)
        dropped = np.abs((a1 @ b1) / (sa * sb)).max() / np.abs(ref).max()
        assert dropped < 2.0 ** -22
