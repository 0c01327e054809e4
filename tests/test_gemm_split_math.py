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
