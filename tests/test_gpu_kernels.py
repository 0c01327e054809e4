"""GPU parity tests of the individual kernels, called through the C ABI (ctypes wrappers in
recsys_benchmark_b200.functional) and compared with the numpy oracle on the same seeded
inputs.  Integer / index work must be bit-exact; fp32 within rtol 1e-5 and a scale-aware
atol (tests/helpers.py)."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def RF():
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200.functional as f

    return f


@pytest.fixture(scope="module")
def L():
    from recsys_benchmark_b200 import _lib

    return _lib


def _keys(n, n_rows, dist, seed):
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        k = rng.integers(0, n_rows, n)
    elif dist == "dup":
        k = rng.integers(0, min(n_rows, 5), n)
    elif dist == "zipf":
        k = np.minimum(rng.zipf(1.05, n) - 1, n_rows - 1)
    elif dist == "sorted_runs":
        k = np.sort(rng.integers(0, max(n_rows // 50, 1), n))
    else:
        raise ValueError(dist)
    return k.astype(np.int64)


# ------------------------------------------------------------------ sort ---
@pytest.mark.parametrize("n,n_rows,dist", [
    (0, 10, "uniform"), (1, 10, "uniform"), (31, 2, "uniform"), (4096, 1000, "uniform"), (4097, 5, "dup"),
    (79872, 1086810, "uniform"), (79872, 1086810, "zipf"), (300001, 17388960, "uniform"),
    (1 << 20, (1 << 31) + 5, "uniform"), (100000, 1, "uniform"),
])
def test_sort_rows_is_a_stable_sort(RF, n, n_rows, dist):
    keys = _keys(n, n_rows, dist, seed=n + 1)
    skeys, perm = RF.sort_rows(torch.from_numpy(keys).to(DEV), n_rows)
    skeys = skeys.cpu().numpy().view(np.uint32).astype(np.int64)
    perm = perm.cpu().numpy().view(np.uint32).astype(np.int64)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(perm, order)
    np.testing.assert_array_equal(skeys, keys[order])


@pytest.mark.parametrize("div,mod", [(5, 0), (0, 5), (1042, 0), (0, 1042)])
def test_sort_applies_qr_index_math_bit_exact(RF, div, mod):
    ids = _keys(50000, 1086810, "uniform", seed=3)
    i1, i2 = O.qr_indices(ids, div or mod)
    keys = i2 if div else i1
    n_rows = int(keys.max()) + 1
    skeys, perm = RF.sort_rows(torch.from_numpy(ids).to(DEV), n_rows, key_div=div, key_mod=mod)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(perm.cpu().numpy().view(np.uint32).astype(np.int64), order)
    np.testing.assert_array_equal(skeys.cpu().numpy().view(np.uint32).astype(np.int64), keys[order])


@pytest.mark.parametrize("d", [1, 2, 3, 5, 7, 20, 64, 65, 1023, 1024, 1025, 65537, 1000003, 2 ** 31 - 1, 2 ** 31,
                               2 ** 32 - 1])
def test_invariant_divisor_division_is_exact(RF, d):
    """The multiply-high division the kernels use for launch-invariant divisors (QR `//`, `%`, shard split) against
    integer arithmetic, on ids that stress the rounding: multiples of d and their neighbours, 2^k boundaries,
    the 32-bit limit, plus one id past 2^32 (exercises the 64-bit fallback)."""
    rng = np.random.default_rng(d % 1000)
    mult = (rng.integers(0, max(1, (2 ** 32 - 1) // d), 4000) * d).astype(np.int64)
    ids = np.concatenate([mult, mult + 1, np.maximum(mult - 1, 0), rng.integers(0, 2 ** 32, 4000),
                          2 ** np.arange(0, 32, dtype=np.int64), 2 ** np.arange(1, 33, dtype=np.int64) - 1,
                          np.array([0, 2 ** 32 - 1, 2 ** 32 + 12345], dtype=np.int64)])
    for div, mod in ((d, 0), (0, d)):
        keys = ids // d if div else ids % d
        keep = keys < 2 ** 32 - 1                      # the sort's key space (0xffffffff is its sentinel)
        sub_ids, sub_keys = ids[keep], keys[keep]
        skeys, perm = RF.sort_rows(torch.from_numpy(sub_ids).to(DEV), int(sub_keys.max()) + 1, key_div=div,
                                   key_mod=mod)
        order = np.argsort(sub_keys, kind="stable")
        np.testing.assert_array_equal(skeys.cpu().numpy().view(np.uint32).astype(np.int64), sub_keys[order])
        np.testing.assert_array_equal(perm.cpu().numpy().view(np.uint32).astype(np.int64), order)


# ------------------------------------------------- segmented reduce/apply ---
@pytest.mark.parametrize("E", [16, 8, 4, 1, 7, 12, 64, 128, 32])
@pytest.mark.parametrize("dist", ["uniform", "dup", "zipf", "sorted_runs"])
def test_segment_reduce_dense_matches_scatter_add(RF, E, dist):
    n, n_rows = 20011, 3000
    keys = _keys(n, n_rows, dist, seed=E)
    rg = np.random.default_rng(E).standard_normal((n, E)).astype(np.float32)
    out = RF.dense_row_grad(torch.from_numpy(keys).to(DEV), torch.from_numpy(rg).to(DEV), n_rows)
    ref = O.scatter_add_dense(keys, rg.astype(np.float64), n_rows)
    assert_close(out.cpu().numpy(), ref, what=f"dense E={E} {dist}")
    # deterministic: a second run is bit-identical
    out2 = RF.dense_row_grad(torch.from_numpy(keys).to(DEV), torch.from_numpy(rg).to(DEV), n_rows)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("b,f", [(1, 1), (33, 3), (1024, 39), (8192, 11), (65536, 39)])
@pytest.mark.parametrize("dist", ["uniform", "dup", "zipf"])
def test_sorted_fc_grad_is_the_scatter_add_and_is_bit_reproducible(RF, b, f, dist):
    """First-order weight gradient from the row-sorted lookups (rsb_fc_grad_sorted): equals the scatter-add of g_y[b]
    over the lookups of each row (aten embedding_dense_backward of FeaturesLinear, deepfm.py:71-76), untouched rows
    stay exactly zero, and - unlike the atomic rsb_fc_grad - two runs are bit-identical."""
    n_rows = max(4, min(200000, b * f // 3))
    keys = _keys(b * f, n_rows, dist, seed=b + f)
    gy = np.random.default_rng(b).standard_normal(b).astype(np.float32)
    rows = torch.from_numpy(keys).to(DEV).view(b, f)
    g = torch.from_numpy(gy).to(DEV)
    pair = RF.sort_rows(rows, n_rows)
    out = RF.fc_grad(rows, g, b, f, (n_rows, 1), n_rows, pair)
    ref = O.scatter_add_dense(keys, np.repeat(gy.astype(np.float64), f)[:, None], n_rows)
    assert_close(out.cpu().numpy(), ref, what=f"fc_grad_sorted b={b} f={f} {dist}")
    untouched = np.ones(n_rows, bool)
    untouched[keys] = False
    assert float(out.cpu().numpy()[untouched].__abs__().sum()) == 0.0
    for _ in range(3):
        assert torch.equal(out, RF.fc_grad(rows, g, b, f, (n_rows, 1), n_rows, pair))
    # the atomic kernel agrees to rounding
    atomic = RF.fc_grad(rows, g, b, f, (n_rows, 1), n_rows, None)
    assert_close(atomic.cpu().numpy(), ref, what="fc_grad atomic")


@pytest.mark.parametrize("n", [1, 31, 32, 33, 255, 256, 257, 1024, 8191])
def test_segment_reduce_chunk_boundaries(RF, n):
    """Runs that start / end exactly on chunk and warp-tile boundaries, incl. one giant run."""
    for keys in [np.zeros(n, np.int64), np.arange(n, dtype=np.int64) // 32, np.arange(n, dtype=np.int64) // 33,
                 np.arange(n, dtype=np.int64), (np.arange(n, dtype=np.int64) + 16) // 64]:
        n_rows = int(keys.max()) + 1
        rg = np.random.default_rng(n).standard_normal((n, 16)).astype(np.float32)
        out = RF.dense_row_grad(torch.from_numpy(keys).to(DEV), torch.from_numpy(rg).to(DEV), n_rows)
        assert_close(out.cpu().numpy(), O.scatter_add_dense(keys, rg.astype(np.float64), n_rows), what=f"n={n}")


@pytest.mark.parametrize("dist", ["uniform", "zipf", "dup"])
def test_fused_sparse_adam_rows_match_torch_sparse_adam(RF, L, dist):
    n, n_rows, E = 30000, 5000, 16
    rng = np.random.default_rng(5)
    w0 = rng.standard_normal((n_rows, E)).astype(np.float32) * 0.1
    w = torch.from_numpy(w0.copy()).to(DEV)
    m = torch.zeros_like(w)
    v = torch.zeros_like(w)
    wo, mo, vo = w0.copy(), np.zeros_like(w0), np.zeros_like(w0)
    # torch's own SparseAdam on the same data as a second reference
    wt = torch.nn.Parameter(torch.from_numpy(w0.copy()).to(DEV))
    opt = torch.optim.SparseAdam([wt], lr=1e-2)
    for step in range(1, 4):
        keys = _keys(n, n_rows, dist, seed=step)
        rg = (rng.standard_normal((n, E)) * 1e-2).astype(np.float32)
        kt, gt = torch.from_numpy(keys).to(DEV), torch.from_numpy(rg).to(DEV)
        skeys, perm = RF.sort_rows(kt, n_rows)
        RF.segment_reduce_apply(L.APPLY_SPARSE_ADAM, skeys, perm, gt, w, m, v, lr=1e-2, step=step)
        uniq, sums = O.coalesce_rows(keys, rg)
        O.sparse_adam_rows(wo, mo, vo, step, uniq, sums, lr=1e-2)
        wt.grad = torch.sparse_coo_tensor(kt.view(1, -1), gt, (n_rows, E))
        opt.step()
        assert_close(w.cpu().numpy(), wo, what=f"w step {step}", atol_scale=2e-5)
        assert_close(m.cpu().numpy(), mo, what=f"m step {step}", atol_scale=2e-5)
        assert_close(v.cpu().numpy(), vo, what=f"v step {step}", atol_scale=2e-5)
        assert_close(w.cpu().numpy(), wt.detach().cpu().numpy(), what=f"w vs torch step {step}", atol_scale=2e-5)
    untouched = np.setdiff1d(np.arange(n_rows), np.unique(np.concatenate([_keys(n, n_rows, dist, s) for s in (1, 2, 3)])))
    np.testing.assert_array_equal(w.cpu().numpy()[untouched], w0[untouched])


def test_fused_sparse_sgd_rows(RF, L):
    n, n_rows, E = 10000, 777, 8
    rng = np.random.default_rng(9)
    w0 = rng.standard_normal((n_rows, E)).astype(np.float32)
    keys = _keys(n, n_rows, "zipf", 2)
    rg = rng.standard_normal((n, E)).astype(np.float32)
    w = torch.from_numpy(w0.copy()).to(DEV)
    skeys, perm = RF.sort_rows(torch.from_numpy(keys).to(DEV), n_rows)
    RF.segment_reduce_apply(L.APPLY_SPARSE_SGD, skeys, perm, torch.from_numpy(rg).to(DEV), w, lr=0.1)
    wo = w0.copy()
    O.sparse_sgd_rows(wo, *O.coalesce_rows(keys, rg), lr=0.1)
    assert_close(w.cpu().numpy(), wo, what="sgd")


@pytest.mark.parametrize("n_rows,E,mod", [(2, 16, 2), (5, 16, 5), (20, 16, 20), (1042, 16, 1042), (7, 3, 7), (3, 8, 0)])
def test_small_table_grad(RF, n_rows, E, mod):
    n = 50000
    ids = _keys(n, 1086810 if mod else n_rows, "uniform", 4)
    keys = ids % mod if mod else ids
    rg = np.random.default_rng(1).standard_normal((n, E)).astype(np.float32)
    out = RF.small_table_grad(torch.from_numpy(ids).to(DEV), torch.from_numpy(rg).to(DEV), n_rows, key_mod=mod)
    assert out is not None
    ref = O.scatter_add_dense(keys, rg.astype(np.float64), n_rows)
    assert_close(out.cpu().numpy(), ref, what="small table", atol_scale=3e-5)
    assert RF.small_table_grad(torch.from_numpy(ids).to(DEV), torch.from_numpy(rg).to(DEV), 10 ** 6) is None


# ------------------------------------------------------ full-table helpers ---
@pytest.mark.parametrize("tt", ["feature_dim", "feature", "dimension", "global"])
def test_pep_threshold_table_and_count(RF, L, tt):
    n, d = 4001, 16
    rng = np.random.default_rng(2)
    w = rng.uniform(-0.5, 0.5, (n, d)).astype(np.float32)
    s = (-1.5 + rng.standard_normal(O.pep_threshold_shape(tt, n, d))).astype(np.float32)
    s_t = torch.from_numpy(s).to(DEV)
    out, cnt = RF.pep_threshold_table(torch.from_numpy(w).to(DEV), s_t, L.PEP_TYPES[tt],
                                      want_out=True, want_count=True)
    # sigmoid(s) in the reference's arithmetic on this device (torch.sigmoid on CUDA): with it the thresholded table,
    # its zero pattern and the non-zero count are the reference's bit for bit
    ref = O.pep_soft_threshold(w, s, sig=torch.sigmoid(s_t).cpu().numpy())
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    assert int(cnt.item()) == int(np.count_nonzero(ref))


@pytest.mark.parametrize("norm", [1, 2])
def test_optembed_eval_weight(RF, norm):
    n, d = 3001, 16
    rng = np.random.default_rng(3)
    w = rng.uniform(-0.2, 0.2, (n, d)).astype(np.float32)
    t = rng.uniform(0.3, 2.5 if norm == 1 else 0.6, n).astype(np.float32)
    k = rng.integers(0, d, n)
    out, cnt = RF.optembed_eval_weight(torch.from_numpy(w).to(DEV), torch.from_numpy(t).to(DEV),
                                       torch.from_numpy(k).to(DEV), norm, want_out=True, want_count=True)
    ref = O.optembed_eval_weight(w, t, None, k, mode_d="feature", norm=norm)
    got = out.cpu().numpy()
    # mask-D (integer index) is exact; mask-E compares a 16-term fp32 sum with t: a row may differ from the oracle only
    # if its norm is within summation-order rounding of the threshold (checked against the fp64 norm)
    bad_rows = np.unique(np.nonzero(got != ref)[0])
    nrm64 = O._row_norm(w.astype(np.float64), norm)
    assert np.all(np.abs(nrm64[bad_rows] - t[bad_rows]) < 4e-7 * np.maximum(nrm64[bad_rows], 1.0))
    flips = sum(int(np.count_nonzero(ref[r])) for r in bad_rows)
    assert abs(int(cnt.item()) - int(np.count_nonzero(ref))) <= flips


def test_mask_table(RF):
    w = torch.randn(1000, 16, device=DEV)
    m = torch.rand(1000, 16, device=DEV) > 0.5
    assert torch.equal(RF.mask_table(w, m), w * m)
