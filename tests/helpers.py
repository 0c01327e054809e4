"""Shared test helpers: golden loading and tolerance rules."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp32 parity rule from SURVEY.md section 8c / BASELINE.md: rtol 1e-5 with a
# scale-aware atol = 1e-5 * max|ref| (element-wise relative error is meaningless
# on cancelling elements of S^2 - Q and of the gradients).
RTOL = 1e-5


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def assert_close(got, ref, rtol=RTOL, atol_scale=1e-5, what="", atol_floor=1e-30):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    atol = max(atol_scale * scale, atol_floor)
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    bound = atol + rtol * np.abs(ref.astype(np.float64))
    bad = err > bound
    assert not bad.any(), (
        f"{what}: {int(bad.sum())}/{bad.size} elements out of tolerance; "
        f"max err {err.max():.3e} at scale {scale:.3e} (atol {atol:.3e}, rtol {rtol})")
