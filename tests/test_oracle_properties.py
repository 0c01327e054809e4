"""Property tests (hypothesis) of the CPU oracle's integer / index functions: these are the functions the CUDA
path must match bit for bit, so their own invariants are pinned independently of the golden files."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import ctr_oracle as O

SET = settings(max_examples=60, deadline=None)


@SET
@given(st.lists(st.integers(1, 5000), min_size=1, max_size=12), st.integers(0, 2 ** 31))
def test_field_offsets_and_global_rows(dims, seed):
    dims = np.asarray(dims, dtype=np.int64)
    off = O.field_offsets(dims)
    assert off[0] == 0 and np.all(np.diff(off) == dims[:-1])
    rng = np.random.default_rng(seed)
    x = np.stack([rng.integers(0, d, 17) for d in dims], 1)
    for dtype in (np.int32, np.int64):
        rows = O.add_offsets(x.astype(dtype), off)
        assert rows.dtype == np.int64 and rows.min() >= 0 and rows.max() < dims.sum()
        assert np.all((rows - off) == x)                      # int32 ids promote, never wrap


@SET
@given(st.integers(1, 2 ** 20), st.integers(1, 5000), st.integers(0, 2 ** 31))
def test_qr_indices_reconstruct_the_id(n, divider, seed):
    ids = np.random.default_rng(seed).integers(0, n, 200)
    i1, i2 = O.qr_indices(ids, divider)
    assert np.all(i2 * divider + i1 == ids) and np.all((0 <= i1) & (i1 < divider))
    assert i2.max() <= (n - 1) // divider                    # emb2 has (N-1)//divider + 1 rows (qr_embedding.py:50-63)


@SET
@given(st.integers(1, 300), st.integers(1, 12), st.integers(0, 2 ** 31))
def test_scatter_add_and_coalesce_preserve_sums(n_rows, e, seed):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, n_rows, 500)
    g = rng.standard_normal((500, e))
    dense = O.scatter_add_dense(rows, g, n_rows)
    np.testing.assert_allclose(dense.sum(0), g.sum(0), rtol=1e-9, atol=1e-9)
    assert np.all(dense[np.setdiff1d(np.arange(n_rows), rows)] == 0)      # untouched rows: exactly zero
    uniq, summed = O.coalesce_rows(rows, g)
    assert np.all(np.diff(uniq) > 0)
    np.testing.assert_allclose(summed, dense[uniq], rtol=1e-9, atol=1e-9)


@SET
@given(st.integers(1, 60), st.integers(1, 32), st.floats(0.0, 1.0), st.integers(0, 2 ** 31))
def test_csr_lookup_equals_dense_gather(n, d, keep, seed):
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal((n, d)) * (rng.random((n, d)) < keep)).astype(np.float32)
    vals, crow, col = O.csr_from_dense(w)
    assert crow[0] == 0 and crow[-1] == len(vals) == int((w != 0).sum())
    for r in range(n):
        seg = col[crow[r]:crow[r + 1]]
        assert np.all(np.diff(seg) > 0)                      # sorted, unique columns: what the kernel relies on
    ids = rng.integers(0, n, (7, 3))
    np.testing.assert_array_equal(O.csr_lookup(vals, crow, col, ids, d), w[ids])


@SET
@given(st.integers(0, 2 ** 40), st.integers(0, 2 ** 33), st.integers(-10 ** 9, 10 ** 9).filter(lambda v: v != 0),
       st.integers(-10 ** 9, 10 ** 9).filter(lambda v: v != 0), st.integers(0, 74517))
def test_dhe_hash_matches_python_integer_arithmetic(item, prefix, slope, bias, prime_idx):
    """int64 with two's-complement wrap-around (what torch does), Python's sign convention for %."""
    p = int(O.first_primes_above(10 ** 6, 74518)[prime_idx])
    m = 1_000_000
    x = slope * (item + prefix + 1) + bias
    x = (x + 2 ** 63) % 2 ** 64 - 2 ** 63                    # wrap to int64
    h = x % p % m                                            # Python: result has the sign of the divisor
    want = np.float32(np.float32(h) / np.float32(m - 1)) * np.float32(2) - np.float32(1)
    with np.errstate(over="ignore"):
        got = O.dhe_universal_hash(np.asarray([item]), prefix, np.asarray([slope]), np.asarray([bias]),
                                   np.asarray([p]), m)[0, 0]
    assert got == want and -1.0 <= got <= 1.0


@SET
@given(st.integers(2, 64), st.integers(0, 2 ** 31))
def test_tril_mask_and_binary_step(d, seed):
    mask = O.tril_mask(d)
    assert mask.shape == (d, d) and np.all(mask.sum(1) == np.arange(1, d + 1))     # row k keeps dims 0..k
    z = np.random.default_rng(seed).standard_normal(100).astype(np.float32) * 2
    step = O.binary_step(z)
    assert set(np.unique(step)) <= {0.0, 1.0} and np.all(step == (z > 0))
    gz = O.binary_step_grad(z)
    assert np.all(gz[np.abs(z) > 1] == 0) and np.all(gz >= 0)
