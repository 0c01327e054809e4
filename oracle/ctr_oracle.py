"""CPU oracle for the CTR hot path of chenxing1999/recsys-benchmark.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`recsys-benchmark_b200/`) may import this file; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and there only as the checker / the timed CPU baseline.

It is a plain numpy restatement (forward AND hand-derived backward) of the
reference's algorithm for the path named in BASELINE.json, every function
citing the reference file:line it follows (paths relative to /root/reference).
The arithmetic that lives in the reference's third-party dependency (PyTorch:
`F.embedding`, `nn.EmbeddingBag`, `torch.optim.SparseAdam/SGD/Adam`,
`einops.einsum`; `torch` is unpinned in the reference's pyproject.toml:17,
torch 2.11.0 is what is installed here) is restated from its published
semantics.

Parity pinning: the reference ships no golden vectors for this path
(SURVEY.md section 8c), so this oracle is pinned against outputs of the
reference itself: `tests/golden/make_golden.py` imports the unmodified
reference from /root/reference and stores inputs, state dicts, logits,
gradients and post-step weights in `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against them.

All functions are dtype-parametric: pass float32 arrays to mimic the
reference's fp32 arithmetic, float64 arrays for a high-precision twin.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# a1. offsets (src/models/deepfm.py:71-76,88 ; src/models/dcn.py:68-74,84)
# --------------------------------------------------------------------------


def field_offsets(field_dims: Sequence[int]) -> np.ndarray:
    """offsets = cumsum([0] + field_dims[:-1]) as int64 [F]."""
    dims = np.asarray(list(field_dims), dtype=np.int64)
    return np.concatenate([np.zeros(1, np.int64), np.cumsum(dims)[:-1]])


def add_offsets(x: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """`x = x + self.offsets`; int32 input promotes to int64 (deepfm.py:88)."""
    return x.astype(np.int64) + offsets.astype(np.int64)[None, :]


# --------------------------------------------------------------------------
# a2. vanilla gather (src/models/embeddings/base.py:53-57,75 -> F.embedding)
# --------------------------------------------------------------------------


def gather_rows(table: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """F.embedding: out[..., :] = table[rows[...], :]; raises on out of range."""
    if rows.size and (rows.min() < 0 or rows.max() >= table.shape[0]):
        raise IndexError("index out of range in self")
    return table[rows]


def scatter_add_dense(rows: np.ndarray, grads: np.ndarray, n_rows: int) -> np.ndarray:
    """embedding_dense_backward: dense zero-filled [N,D] grad, duplicates summed."""
    out = np.zeros((n_rows,) + grads.shape[rows.ndim:], dtype=grads.dtype)
    np.add.at(out, rows.reshape(-1), grads.reshape((-1,) + grads.shape[rows.ndim:]))
    return out


def coalesce_rows(rows: np.ndarray, grads: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Sparse-COO coalesce (what SparseAdam does first, torch/optim/_functional.py:43):
    unique sorted rows and the per-row sums, summed in original (stable) order."""
    flat = rows.reshape(-1)
    g = grads.reshape(flat.shape[0], -1)
    order = np.argsort(flat, kind="stable")
    srows = flat[order]
    if srows.size == 0:
        return srows, g[:0]
    heads = np.concatenate([[True], srows[1:] != srows[:-1]])
    uniq = srows[heads]
    seg = np.cumsum(heads) - 1
    sums = np.zeros((uniq.shape[0], g.shape[1]), dtype=g.dtype)
    np.add.at(sums, seg, g[order])
    return uniq, sums


# --------------------------------------------------------------------------
# a3/a4. DeepFM first order + FM second order (src/models/deepfm.py:49,91-98)
# --------------------------------------------------------------------------


def first_order(fc: np.ndarray, bias: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """EmbeddingBag(N,1,'sum')(x) + _bias -> [B,1] (deepfm.py:49,51,95)."""
    return fc[rows, 0].sum(axis=1, keepdims=True) + bias.reshape(1, 1)


def fm_second_order(emb: np.ndarray) -> np.ndarray:
    """0.5 * sum_d((sum_f e)^2 - sum_f e^2) -> [B,1] (deepfm.py:91-92,98)."""
    square_of_sum = emb.sum(axis=1) ** 2
    sum_of_square = (emb ** 2).sum(axis=1)
    half = emb.dtype.type(0.5)
    return half * (square_of_sum - sum_of_square).sum(axis=1, keepdims=True)


def deepfm_yfm(emb, fc, bias, rows) -> np.ndarray:
    """y_fm = first order + second order (deepfm.py:95-98) -> [B,1]."""
    return first_order(fc, bias, rows) + fm_second_order(emb)


def fm_backward(emb: np.ndarray, g_yfm: np.ndarray, g_deep: Optional[np.ndarray]) -> np.ndarray:
    """d loss / d emb[b,f,:] = g_y[b]*(S_b - e[b,f,:]) + g_deep[b,f,:].

    g_yfm: [B] or [B,1]; g_deep: [B,F*D] (grad arriving through
    `emb.reshape(b, F*D)` from the deep branch, deepfm.py:100-102) or None."""
    b, f, d = emb.shape
    s = emb.sum(axis=1, keepdims=True)
    g = g_yfm.reshape(b, 1, 1) * (s - emb)
    if g_deep is not None:
        g = g + g_deep.reshape(b, f, d)
    return g


def first_order_backward(rows: np.ndarray, g_yfm: np.ndarray, n_rows: int):
    """Dense [N,1] grad of fc.weight and scalar grad of _bias."""
    b, f = rows.shape
    g = np.broadcast_to(g_yfm.reshape(b, 1), (b, f)).astype(g_yfm.dtype)
    g_fc = np.zeros((n_rows, 1), dtype=g_yfm.dtype)
    np.add.at(g_fc[:, 0], rows.reshape(-1), g.reshape(-1))
    return g_fc, g_yfm.sum().reshape(1)


# --------------------------------------------------------------------------
# a6. QR hashing (src/models/embeddings/qr_embedding.py:10-113)
# --------------------------------------------------------------------------


def qr_default_divider(num_item: int) -> int:
    """qr_embedding.py:47-48."""
    return int(math.sqrt(num_item))


def qr_table_sizes(num_item: int, divider: int, hidden: int, operation: str):
    """(rows1, rows2, emb_size): qr_embedding.py:50-63."""
    emb = hidden // 2 if operation == "cat" else hidden
    return divider, (num_item - 1) // divider + 1, emb


def qr_indices(rows: np.ndarray, divider: int) -> Tuple[np.ndarray, np.ndarray]:
    """i1 = x % divider, i2 = x // divider on non-negative int64 (qr_embedding.py:96-97).
    Bit-exact requirement."""
    r = rows.astype(np.int64)
    return r % np.int64(divider), r // np.int64(divider)


def qr_forward(emb1, emb2, rows, divider: int, operation: str = "mult"):
    """qr_embedding.py:95-109.  `cat` concatenates on dim=1 (the FIELD axis for a
    [B,F] input -> [B,2F,D/2]; for 1-D ids -> [B,D]) exactly like the reference."""
    i1, i2 = qr_indices(rows, divider)
    e1, e2 = emb1[i1], emb2[i2]
    if operation == "cat":
        return np.concatenate([e1, e2], axis=1)
    if operation == "add":
        return e1 + e2
    if operation == "mult":
        return e1 * e2
    raise NotImplementedError(operation)


def qr_backward(emb1, emb2, rows, divider: int, operation: str, g_out):
    """Dense grads (g_emb1, g_emb2) of the QR composition for `g_out` shaped like
    the forward output."""
    i1, i2 = qr_indices(rows, divider)
    if operation == "cat":
        nf = rows.shape[1] if rows.ndim == 2 else None
        if nf is not None:
            g1, g2 = g_out[:, :nf], g_out[:, nf:]
        else:
            e = emb1.shape[1]
            g1, g2 = g_out[:, :e], g_out[:, e:]
    elif operation == "add":
        g1, g2 = g_out, g_out
    elif operation == "mult":
        g1, g2 = g_out * emb2[i2], g_out * emb1[i1]
    else:
        raise NotImplementedError(operation)
    return (scatter_add_dense(i1, np.ascontiguousarray(g1), emb1.shape[0]),
            scatter_add_dense(i2, np.ascontiguousarray(g2), emb2.shape[0]))


# --------------------------------------------------------------------------
# a7/a8. PEP (src/models/embeddings/pep_embedding.py:12-229)
# --------------------------------------------------------------------------


def sigmoid(s: np.ndarray) -> np.ndarray:
    one = s.dtype.type(1)
    with np.errstate(over="ignore"):
        return one / (one + np.exp(-s))


def pep_soft_threshold(v: np.ndarray, s: np.ndarray, sig: Optional[np.ndarray] = None) -> np.ndarray:
    """sign(v) * relu(abs(v) - sigmoid(s)) (pep_embedding.py:91-92); s broadcasts.

    `sig`: sigmoid(s) precomputed by the caller in the arithmetic of the device the reference runs on
    (torch.sigmoid on CUDA rounds differently from numpy's exp in the last bit); with it every
    `abs(v) > sigmoid(s)` decision is the reference's own, bit for bit."""
    sg = sigmoid(s) if sig is None else sig
    return np.sign(v) * np.maximum(np.abs(v) - sg, v.dtype.type(0))


def pep_threshold_shape(threshold_type: str, num_item: int, hidden: int):
    """pep_embedding.py:94-117."""
    return {"global": (1,), "dimension": (hidden,), "feature": (num_item, 1),
            "feature_dim": (num_item, hidden)}[threshold_type]


def pep_forward(weight, s, rows, sig: Optional[np.ndarray] = None) -> np.ndarray:
    """F.embedding(x, soft_threshold(weight, s)) (pep_embedding.py:82-89)."""
    return gather_rows(pep_soft_threshold(weight, s, sig), rows)


def pep_backward(weight, s, rows, g_out, sig: Optional[np.ndarray] = None):
    """Dense grads (g_weight[N,D], g_s[shape of s]).

    d/dv = 1[abs(v) > sigmoid(s)]; d/ds = -sign(v) 1[...] sigmoid(s)(1-sigmoid(s)),
    reduced to the broadcast shape of `s`."""
    n, d = weight.shape
    g_table = scatter_add_dense(rows, g_out, n)  # grad wrt thresholded table
    sg = np.broadcast_to(sigmoid(s) if sig is None else sig, weight.shape)
    keep = (np.abs(weight) - sg) > 0
    g_w = np.where(keep, np.sign(weight) ** 2 * g_table, 0).astype(weight.dtype)
    g_s_full = np.where(keep, -np.sign(weight) * g_table * sg * (1 - sg), 0).astype(weight.dtype)
    if s.shape == (1,):
        g_s = g_s_full.sum().reshape(1)
    elif s.shape == (d,):
        g_s = g_s_full.sum(axis=0)
    elif s.shape == (n, 1):
        g_s = g_s_full.sum(axis=1, keepdims=True)
    else:
        g_s = g_s_full
    return g_w, g_s.astype(weight.dtype)


def pep_count_nonzero(weight, s, sig: Optional[np.ndarray] = None) -> int:
    """get_num_params: count_nonzero(soft_threshold(weight, s)) (pep_embedding.py:127-130)."""
    return int(np.count_nonzero(pep_soft_threshold(weight, s, sig)))


def pep_retrain_mask(weight_final, s_final, sig: Optional[np.ndarray] = None) -> np.ndarray:
    """mask = (abs(w) - sigmoid(s)) > 0 (pep_embedding.py:203)."""
    return (np.abs(weight_final) - (sigmoid(s_final) if sig is None else sig)) > 0


def masked_forward(weight, mask, rows) -> np.ndarray:
    """F.embedding(x, weight * mask) (pep_embedding.py:215-221; deepfm_opt_embed.py:693-706)."""
    return gather_rows(weight * mask.astype(weight.dtype), rows)


def masked_backward(weight, mask, rows, g_out) -> np.ndarray:
    return scatter_add_dense(rows, g_out, weight.shape[0]) * mask.astype(weight.dtype)


# --------------------------------------------------------------------------
# f-1. CERP (src/models/embeddings/cerp_embedding.py:142-207)
# --------------------------------------------------------------------------


def cerp_indices(rows: np.ndarray, num_item: int, bucket_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """q_idx = trunc(x / ceil(num_item / bucket)), p_idx = x % bucket (cerp_embedding.py:69,145-146)."""
    per_row = int(np.ceil(num_item / bucket_size))
    r = rows.astype(np.int64)
    return r // np.int64(per_row), r % np.int64(bucket_size)


def cerp_forward(p_w, q_w, p_t, q_t, rows, num_item: int) -> np.ndarray:
    """emb = soft(Q)[q_idx] + soft(P)[p_idx] (cerp_embedding.py:134-153)."""
    qi, pi = cerp_indices(rows, num_item, p_w.shape[0])
    return pep_soft_threshold(q_w, q_t)[qi] + pep_soft_threshold(p_w, p_t)[pi]


def cerp_backward(p_w, q_w, p_t, q_t, rows, num_item: int, g_out):
    """Returns dense (g_p_w, g_p_t, g_q_w, g_q_t)."""
    qi, pi = cerp_indices(rows, num_item, p_w.shape[0])
    n = p_w.shape[0]
    res = []
    for w, t, idx in ((p_w, p_t, pi), (q_w, q_t, qi)):
        g_table = scatter_add_dense(idx, g_out, n)
        sg = sigmoid(t)
        keep = (np.abs(w) - sg) > 0
        res.append(np.where(keep, np.sign(w) ** 2 * g_table, 0).astype(w.dtype))
        res.append(np.where(keep, -np.sign(w) * g_table * sg * (1 - sg), 0).astype(w.dtype))
    return tuple(res)


def cerp_retrain_masks(target: Dict[str, np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """(q_mask, p_mask) = |w| - sigmoid(threshold) > 0 of the searched checkpoint
    (RetrainCerpEmbedding.load_mask, cerp_embedding.py:302-318)."""
    return tuple((np.abs(target[w]) - sigmoid(target[t])) > 0
                 for w, t in (("q_weight", "q_threshold"), ("p_weight", "p_threshold")))


def cerp_retrain_forward(p_w, q_w, p_mask, q_mask, rows, num_item: int) -> np.ndarray:
    """emb = (Q * q_mask)[q_idx] + (P * p_mask)[p_idx] (RetrainCerpEmbedding.forward, cerp_embedding.py:329-349)."""
    qi, pi = cerp_indices(rows, num_item, p_w.shape[0])
    return (q_w * q_mask)[qi] + (p_w * p_mask)[pi]


def cerp_retrain_backward(p_mask, q_mask, rows, num_item: int, g_out):
    """Dense (g_p_weight, g_q_weight): scatter-add of g_out by index, times the fixed mask."""
    qi, pi = cerp_indices(rows, num_item, p_mask.shape[0])
    n = p_mask.shape[0]
    return scatter_add_dense(pi, g_out, n) * p_mask, scatter_add_dense(qi, g_out, n) * q_mask


def cerp_prune_loss(p_w, q_w, p_t, q_t, K=100):
    """get_prune_loss (cerp_embedding.py:205-207)."""
    emb = pep_soft_threshold(p_w, p_t) + pep_soft_threshold(q_w, q_t)
    return -np.sum(np.tanh(emb * K) ** 2)


# --------------------------------------------------------------------------
# a9/a10. OptEmbed (deepfm_opt_embed.py:40-307,633-718 ; optembed_utils.py)
# --------------------------------------------------------------------------


def tril_mask(hidden: int) -> np.ndarray:
    """get_mask: matrix[i][j] = 1 if i >= j (optembed_utils.py:10-22)."""
    return np.tril(np.ones((hidden, hidden), dtype=bool))


def binary_step(z: np.ndarray) -> np.ndarray:
    """BinaryStep.forward: (inp > 0).float() (optembed_utils.py:30-33)."""
    return (z > 0).astype(z.dtype)


def binary_step_grad(z: np.ndarray) -> np.ndarray:
    """BinaryStep.backward surrogate (optembed_utils.py:35-44):
    |z|>1 -> 0 ; 0.4<|z|<=1 -> 0.4 ; else 2-4|z|."""
    a = np.abs(z)
    add = 2 - 4 * a
    add = np.where(a > 1, 0.0, add)
    add = np.where((a <= 1) & (a > 0.4), 0.4, add)
    return add.astype(z.dtype)


def _row_norm(e: np.ndarray, norm: int) -> np.ndarray:
    if norm == 1:
        return np.abs(e).sum(axis=-1)
    return np.sqrt((e * e).sum(axis=-1))


def optembed_train_forward(weight, t_param, rows, mask_d_idx, norm: int = 1):
    """Supernet training forward (deepfm_opt_embed.py:219-226, optembed_utils.py:101-104).

    rows [B,F] global ids, t_param [F] (field mode) or None (mask-E disabled,
    `deepfm_optembed_d`), mask_d_idx [B,F] = the torch.randint(0,D) draw."""
    e = gather_rows(weight, rows)
    if t_param is not None:
        z = _row_norm(e, norm) - t_param[None, :]
        e = e * binary_step(z)[..., None]
    d = weight.shape[1]
    md = tril_mask(d)[mask_d_idx].astype(weight.dtype)
    return md * e


def optembed_train_backward(weight, t_param, rows, mask_d_idx, g_out, norm: int = 1):
    """Returns (g_weight dense [N,D], g_t [F] or None)."""
    n, d = weight.shape
    e = gather_rows(weight, rows)
    md = tril_mask(d)[mask_d_idx].astype(weight.dtype)
    u = g_out * md  # grad wrt (e * mask_e)
    if t_param is None:
        return scatter_add_dense(rows, u, n), None
    nrm = _row_norm(e, norm)
    z = nrm - t_param[None, :]
    me = binary_step(z)
    g_e = u * me[..., None]
    g_me = (u * e).sum(axis=-1)
    g_z = g_me * binary_step_grad(z)
    if norm == 1:
        dn = np.sign(e)
    else:
        with np.errstate(invalid="ignore", divide="ignore"):
            dn = np.where(nrm[..., None] > 0, e / nrm[..., None], 0)
    g_e = g_e + g_z[..., None] * dn
    g_t = -g_z.sum(axis=0)
    return scatter_add_dense(rows, g_e.astype(weight.dtype), n), g_t.astype(weight.dtype)


def optembed_feature_threshold(t_param, field_dims) -> np.ndarray:
    """_transform_t_to_feat: repeat_interleave(t, field_dims) (optembed_utils.py:76-86)."""
    return np.repeat(t_param, np.asarray(field_dims, dtype=np.int64))


def optembed_eval_weight(weight, t_param, field_dims, mask_d_idx=None, mode_d="field", norm=1):
    """get_weight in eval mode (deepfm_opt_embed.py:148-202): full-table mask-E
    (optembed_utils.py:88-99), optional mask-D given as indices per field/feature."""
    emb = weight
    if t_param is not None:
        t = optembed_feature_threshold(t_param, field_dims) if t_param.shape[0] != weight.shape[0] else t_param
        emb = weight * binary_step(_row_norm(weight, norm) - t)[:, None]
    if mask_d_idx is not None:
        idx = np.asarray(mask_d_idx, dtype=np.int64)
        if mode_d == "field":
            idx = np.repeat(idx, np.asarray(field_dims, dtype=np.int64))
        emb = emb * tril_mask(weight.shape[1])[idx].astype(weight.dtype)
    return emb


def optembed_l_s(t_param) -> np.ndarray:
    """get_l_s = exp(-t).sum() (deepfm_opt_embed.py:143-146)."""
    return np.exp(-t_param).sum()


def retrain_optembed_mask(mask_e, mask_d_idx, field_dims, hidden, mode_d="field") -> np.ndarray:
    """RetrainOptEmbed.init_mask (deepfm_opt_embed.py:666-691) -> bool/int [N,D]."""
    idx = np.asarray(mask_d_idx, dtype=np.int64)
    if mode_d == "field":
        idx = np.repeat(idx, np.asarray(field_dims, dtype=np.int64))
    return tril_mask(hidden)[idx] * np.asarray(mask_e)[:, None]


# --------------------------------------------------------------------------
# a14. row updates (torch.optim.SparseAdam / SGD / Adam as used by
#      src/models/deepfm.py:155-219)
# --------------------------------------------------------------------------


def sparse_adam_rows(w, m, v, step: int, uniq_rows, g_sums, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch/optim/_functional.py:24-84 on coalesced rows; in place on w, m, v.
    `step` is the 1-based global step counter."""
    f = w.dtype.type
    g = g_sums.astype(w.dtype)
    old_m = m[uniq_rows]
    old_v = v[uniq_rows]
    m_upd = (g - old_m) * f(1 - beta1)
    v_upd = (g * g - old_v) * f(1 - beta2)
    m[uniq_rows] = old_m + m_upd
    v[uniq_rows] = old_v + v_upd
    numer = m_upd + old_m
    denom = np.sqrt(v_upd + old_v) + f(eps)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr * math.sqrt(bc2) / bc1
    w[uniq_rows] = w[uniq_rows] + f(-step_size) * (numer / denom)


def sparse_sgd_rows(w, uniq_rows, g_sums, lr):
    """torch.optim.SGD on a sparse grad, no momentum, wd=0 (deepfm.py:203-216): p += -lr*g."""
    w[uniq_rows] = w[uniq_rows] + w.dtype.type(-lr) * g_sums.astype(w.dtype)


def dense_adam(w, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam (coupled L2 weight decay), single tensor; in place."""
    f = w.dtype.type
    if weight_decay:
        g = g + f(weight_decay) * w
    m[...] = m + (g - m) * f(1 - beta1)  # lerp form used by torch
    v[...] = v * f(beta2) + (g * g) * f(1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    w[...] = w - f(lr / bc1) * (m / denom)


# --------------------------------------------------------------------------
# a11. DCN-Mix cross head (src/models/layer_dcn.py:8-24,90-115)
# --------------------------------------------------------------------------


def dcn_mix_layer_forward(x0, xl, V, C, U, bias, gates, softmax_gate=False):
    """One layer of DCN_MixHead.forward (layer_dcn.py:94-113).

    x0, xl [B,Dm]; V [E,Dm,r]; C [E,r,r]; U [E,r,Dm]; bias [1,Dm]; gates [E,Dm,1].
    Returns (x_next, cache) ."""
    h1p = np.einsum("bd,edr->ber", xl, V)           # x @ V, permuted (layer_dcn.py:20-21)
    h1 = np.tanh(h1p)
    h2p = np.einsum("ber,ers->bes", h1, C)          # :22
    h2 = np.tanh(h2p)
    eo = np.einsum("ber,erd->bed", h2, U)           # :23
    ee = (eo + bias[None, :, :]) * x0[:, None, :]   # :102-103
    g = np.einsum("bd,ed->be", xl, gates[:, :, 0])  # :107-109
    graw = g
    if softmax_gate:
        g = np.exp(g - g.max(axis=1, keepdims=True))
        g = g / g.sum(axis=1, keepdims=True)
    x_next = np.einsum("be,bed->bd", g, ee) + xl     # :113
    return x_next, dict(h1=h1, h2=h2, eo=eo, ee=ee, g=g, graw=graw, xl=xl)


def dcn_mix_forward(x0, params: Dict[str, List[np.ndarray]], softmax_gate=False):
    """DCN_MixHead.forward. params: lists 'V','C','U','biases' per layer + 'gates'."""
    xl = x0
    caches = []
    for l in range(len(params["V"])):
        xl, c = dcn_mix_layer_forward(x0, xl, params["V"][l], params["C"][l], params["U"][l],
                                      params["biases"][l], params["gates"], softmax_gate)
        caches.append(c)
    return xl, caches


def dcn_mix_backward(x0, params, caches, g_out, softmax_gate=False):
    """Backward of dcn_mix_forward.  Returns (g_x0, grads dict like params)."""
    nl = len(params["V"])
    gV, gC, gU, gB = [None] * nl, [None] * nl, [None] * nl, [None] * nl
    g_gates = np.zeros_like(params["gates"])
    g_x0 = np.zeros_like(x0)
    g_xl = g_out
    for l in reversed(range(nl)):
        c = caches[l]
        V, C, U, bias = params["V"][l], params["C"][l], params["U"][l], params["biases"][l]
        xl, h1, h2, eo, ee, g = c["xl"], c["h1"], c["h2"], c["eo"], c["ee"], c["g"]
        g_next = g_xl
        g_xl = g_next.copy()                                  # residual
        g_g = np.einsum("bd,bed->be", g_next, ee)
        g_ee = g[:, :, None] * g_next[:, None, :]
        if softmax_gate:
            g_g = g * (g_g - (g_g * g).sum(axis=1, keepdims=True))
        g_gates[:, :, 0] += np.einsum("be,bd->ed", g_g, xl)
        g_xl += np.einsum("be,ed->bd", g_g, params["gates"][:, :, 0])
        g_x0 += (g_ee * (eo + bias[None])).sum(axis=1)
        g_eo = g_ee * x0[:, None, :]
        gB[l] = g_eo.sum(axis=(0, 1)).reshape(1, -1)
        gU[l] = np.einsum("ber,bed->erd", h2, g_eo)
        g_h2p = np.einsum("bed,erd->ber", g_eo, U) * (1 - h2 * h2)
        gC[l] = np.einsum("ber,bes->ers", h1, g_h2p)
        g_h1p = np.einsum("bes,ers->ber", g_h2p, C) * (1 - h1 * h1)
        gV[l] = np.einsum("bd,ber->edr", xl, g_h1p)
        g_xl += np.einsum("ber,edr->bd", g_h1p, V)
    g_x0 = g_x0 + g_xl                                        # x_l at layer 0 is x_0
    return g_x0, dict(V=gV, C=gC, U=gU, biases=gB, gates=g_gates)


# --------------------------------------------------------------------------
# a5/a12. dense tails in eval mode (only to check whole-model logits; the MLP is
#         off the hot path and stays in cuBLAS on the GPU side)
# --------------------------------------------------------------------------


def mlp_eval(x, layers: List[Dict[str, np.ndarray]]):
    """Sequential(Linear[/BatchNorm1d]/ReLU/Dropout ... Linear) in eval mode
    (src/models/deepfm.py:55-66 ; src/models/dcn.py:56-66)."""
    for i, ly in enumerate(layers):
        x = x @ ly["weight"].T + ly["bias"]
        if i == len(layers) - 1:
            break
        if "bn_weight" in ly:
            x = (x - ly["bn_mean"]) / np.sqrt(ly["bn_var"] + x.dtype.type(1e-5)) * ly["bn_weight"] + ly["bn_bias"]
        x = np.maximum(x, 0)
    return x


def mlp_layers_from_state(state: Dict[str, np.ndarray], prefix: str, dtype=None):
    """Collect Linear / BatchNorm1d tensors of an nn.Sequential from a state dict."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in state if k.startswith(prefix)})
    layers: List[Dict[str, np.ndarray]] = []
    for i in idx:
        w = state.get(f"{prefix}{i}.weight")
        if w is None:
            continue
        if w.ndim == 2:
            layers.append(dict(weight=w, bias=state[f"{prefix}{i}.bias"]))
        else:
            layers[-1].update(bn_weight=w, bn_bias=state[f"{prefix}{i}.bias"],
                              bn_mean=state[f"{prefix}{i}.running_mean"],
                              bn_var=state[f"{prefix}{i}.running_var"])
    if dtype is not None:
        layers = [{k: v.astype(dtype) for k, v in ly.items()} for ly in layers]
    return layers


def deepfm_logits_eval(state: Dict[str, np.ndarray], emb: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """DeepFM.forward after the embedding, eval mode (deepfm.py:91-105) -> [B]."""
    y_fm = deepfm_yfm(emb, state["fc.weight"], state["_bias"], rows)
    b = emb.shape[0]
    deep = mlp_eval(emb.reshape(b, -1), mlp_layers_from_state(state, "_deep_branch."))
    return (y_fm + deep)[:, 0]


def csr_from_dense(weight: np.ndarray):
    """Dense [N,D] -> (values, crow_indices, col_indices) in torch.to_sparse_csr() order
    (PrunedEmbedding.from_weight, pruned_embedding.py:38-49): row-major scan of the non-zeros."""
    nz = weight != 0
    crow = np.zeros(weight.shape[0] + 1, dtype=np.int64)
    np.cumsum(nz.sum(1), out=crow[1:])
    r, c = np.nonzero(nz)
    return weight[r, c].astype(np.float32), crow, c.astype(np.int64)


def csr_lookup(values: np.ndarray, crow: np.ndarray, col: np.ndarray, ids: np.ndarray, d: int) -> np.ndarray:
    """PrunedEmbedding.forward (pruned_embedding.py:89-138) / csr_embedding_lookup(_cpu) (:140-173,186-203):
    out[i, :] = 0; out[i, col[j]] = values[j] for j in [crow[ids[i]], crow[ids[i]+1]).  ids any shape -> [..., d]."""
    flat = ids.reshape(-1).astype(np.int64)
    out = np.zeros((flat.shape[0], d), dtype=np.float32)
    for i, r in enumerate(flat):
        for j in range(int(crow[r]), int(crow[r + 1])):
            out[i, int(col[j])] = values[j]
    return out.reshape(*ids.shape, d)


def dhe_universal_hash(ids: np.ndarray, prefix: int, slopes: np.ndarray, bias: np.ndarray, primes: np.ndarray,
                       m: int = 1_000_000) -> np.ndarray:
    """DHEmbedding._get_universal_hash(_batch) (dh_embedding.py:194-236): ids [...] -> codes [..., k] fp32.
    int64 arithmetic, `%` with the sign of the divisor (numpy == Python == torch.remainder), then the fp32
    true division by (m - 1) and `* 2 - 1`."""
    item = ids.astype(np.int64)[..., None] + np.int64(prefix) + np.int64(1)
    h = (slopes.astype(np.int64) * item + bias.astype(np.int64)) % primes.astype(np.int64) % np.int64(m)
    enc = h.astype(np.float32) / np.float32(m - 1)
    return enc * np.float32(2) - np.float32(1)


def first_primes_above(lo: int, count: int) -> np.ndarray:
    """The prime table DHE draws its moduli from: src/assets/large_prime_74518.json is exactly the first
    74 518 primes above 10^6 (checked against the file in tests/golden/make_golden.py's container)."""
    hi = int(lo * 2.2) + 1000
    while True:
        sieve = np.ones(hi + 1, dtype=bool)
        sieve[:2] = False
        for i in range(2, int(hi ** 0.5) + 1):
            if sieve[i]:
                sieve[i * i::i] = False
        pr = np.nonzero(sieve)[0]
        pr = pr[pr > lo]
        if len(pr) >= count:
            return pr[:count].astype(np.int64)
        hi *= 2


def mish(x):
    """nn.Mish: x * tanh(softplus(x)) (torch thresholds softplus at 20)."""
    sp = np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))
    return x * np.tanh(sp)


def dhe_mlp_eval(x: np.ndarray, state: Dict[str, np.ndarray], prefix: str, use_bn: int) -> np.ndarray:
    """DHEmbedding._seq in eval mode (dh_embedding.py:101-116): per hidden size Linear then
    use_bn == 1: Mish, BatchNorm1d | use_bn == 2: BatchNorm1d, Mish | else: Mish."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in state if k.startswith(prefix)})
    for i in idx:
        w = state.get(f"{prefix}{i}.weight")
        if w is None:
            continue
        if w.ndim == 2:
            x = x @ w.T + state[f"{prefix}{i}.bias"]
            if use_bn != 2:
                x = mish(x)
        else:
            x = (x - state[f"{prefix}{i}.running_mean"]) / np.sqrt(
                state[f"{prefix}{i}.running_var"] + x.dtype.type(1e-5)) * w + state[f"{prefix}{i}.bias"]
            if use_bn == 2:
                x = mish(x)
    return x


def bce_with_logits_grad(logits: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """d mean-BCEWithLogits / d logits = (sigmoid(z) - y) / B (trainer/deepfm.py:33,51)."""
    return (sigmoid(logits) - labels.astype(logits.dtype)) / logits.dtype.type(logits.shape[0])
