"""fp32 matrices held as three bf16 planes, and the hand-written tensor-core GEMM on them (rsb_gemm_planes).

`Planes` is the operand format of csrc/gemm/planes_gemm.cu: X = X0 + X1 + X2 with 8 mantissa bits per plane, written
ONCE per operand (by `split`, or by a producer kernel's epilogue) and read by every GEMM that uses the operand - the
forward GEMM and the weight-gradient GEMM of a layer share the activation planes, the dX and dW GEMMs share the
gradient planes, and the planes of an nn.Linear weight are rebuilt only when the weight has changed."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from . import functional as RF


class Planes:
    """bf16 [3, rows, ld] (ld = cols rounded up to 8) + the logical shape of the fp32 matrix it represents."""

    __slots__ = ("data", "rows", "cols")

    def __init__(self, data: torch.Tensor, rows: int, cols: int):
        self.data, self.rows, self.cols = data, rows, cols

    @property
    def ld(self) -> int:
        return self.data.shape[2]

    def operand(self, mn_major: bool, row_step: int = 0, col_step: int = 0, cols: Optional[int] = None) -> L.PlanesOperand:
        """`cols`: extent of the stored columns the GEMM may read (TMA zero-fills beyond it).  Default: the logical
        width - a forward GEMM must not reduce over the padding / ones column; a weight-gradient GEMM that wants the
        bias gradient passes the padded width + 8."""
        return L.PlanesOperand(self.data.data_ptr(), self.rows, self.cols if cols is None else cols, self.ld,
                               self.data.stride(0), int(mn_major), row_step, col_step)

    def float(self) -> torch.Tensor:
        """The represented fp32 matrix (sum of the planes) - for tests."""
        return self.data.float().sum(0)[:, :self.cols]


def alloc(rows: int, cols: int, device, ones_col: bool = False) -> Planes:
    """Uninitialised planes for a [rows, cols] matrix (the writer fills every column up to ld)."""
    ld = (cols + 7) // 8 * 8 + (8 if ones_col else 0)
    return Planes(torch.empty(3, rows, ld, dtype=torch.bfloat16, device=device), rows, cols)


def split(x: torch.Tensor, transpose: bool = False, ones_col: bool = False) -> Planes:
    """fp32 [rows, cols] (unit column stride) -> Planes of x (or of x^T).  ones_col: a column of ones after the
    8-padded data columns (see `gemm_dw`)."""
    lib = L.load()
    dev = L.require_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    orows, ocols = (cols, rows) if transpose else (rows, cols)
    out = alloc(orows, ocols, dev, ones_col)
    RF._call("split_planes", lib.rsb_split_planes, L.ptr(x), rows, cols, x.stride(0), int(transpose), int(ones_col),
             L.ptr(out.data), out.ld, out.data.stride(0), L.stream_ptr(dev),
             nbytes=rows * cols * 4 + 3 * orows * out.ld * 2)
    return out


def _dropout_epilogue(mode: int, out: Optional[Planes], mask: torch.Tensor, p: float, ones_col: bool) -> L.GemmEpilogue:
    return L.GemmEpilogue(mode, out.data.data_ptr() if out is not None else None, out.ld if out is not None else 0,
                          out.data.stride(0) if out is not None else 0, int(ones_col), mask.data_ptr(), float(p))


def dropout_keep_mask(shape, p: float, seed: int, offset: int, offset_dev: Optional[torch.Tensor], device) -> torch.Tensor:
    """uint8 keep bits (1 with probability 1 - p) of an activation, from the library's Philox stream."""
    lib = L.load()
    mask = torch.empty(shape, dtype=torch.uint8, device=device)
    RF._call("dropout_keep_mask", lib.rsb_dropout_keep_mask, L.ptr(mask), mask.numel(), float(p),
             seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, L.ptr(offset_dev), L.stream_ptr(device),
             nbytes=mask.numel())
    return mask


def gemm(a: Planes, b: Planes, m: int, n: int, k: int, *, a_mn_major: bool = False, b_mn_major: bool = False,
         bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 0.0,
         c: Optional[torch.Tensor] = None, split_k: int = 0, batch: int = 1, a_steps=(0, 0), b_steps=(0, 0),
         d_batch_stride: int = 0, epilogue: Optional[L.GemmEpilogue] = None, want_out: bool = True,
         b_cols: Optional[int] = None):
    """D [m, n] = alpha * A B^T + beta * C + bias on the tensor cores.

    a_mn_major = False: `a` stores [m, k];  True: `a` stores [k, m] (reduction over the stored rows).
    b_mn_major = False: `b` stores [n, k] (nn.Linear weight layout);  True: `b` stores [k, n].
    batch > 1: operand l starts at stored (row, col) + l * steps; D[l] = out.data + l * d_batch_stride floats.
    epilogue: a fused rsb_gemm_epilogue (see `linear_relu_dropout` / `dx_masked`)."""
    lib = L.load()
    dev = a.data.device
    if out is None and want_out:
        out = torch.empty(m, n, dtype=torch.float32, device=dev)
    nb = lib.rsb_gemm_planes_workspace_bytes(m, n, k, batch, split_k)
    ws = RF._ws(nb, dev)
    oa = a.operand(a_mn_major, *a_steps)
    ob = b.operand(b_mn_major, *b_steps, cols=b_cols)
    kind = ("gemm_planes" if epilogue is None else
            ("gemm_planes_relu_dropout", "gemm_planes_relu_dropout", "gemm_planes_masked", "gemm_planes_masked")[epilogue.mode])
    if a_mn_major and b_mn_major:
        kind = "gemm_planes_dw"
    RF._call(kind, lib.rsb_gemm_planes, C.byref(oa), C.byref(ob), m, n, k, batch, split_k, L.ptr(c), L.ptr(out),
             out.stride(0) if out is not None else n, d_batch_stride, L.ptr(bias), alpha, beta,
             C.byref(epilogue) if epilogue is not None else None, L.ptr(ws), ws.numel(), L.stream_ptr(dev),
             nbytes=2 * m * n * k * batch)       # `bytes` slot of the timer carries the fp32-equivalent FLOPs here
    return out


def linear_relu_dropout(xp: Planes, wp: Planes, bias: Optional[torch.Tensor], p: float, seed: int, offset: int,
                        offset_dev: Optional[torch.Tensor] = None, ones_col: bool = True):
    """dropout_p(relu(x W^T + b)) in ONE launch: returns (planes of y [M, N] (+ ones column), keep mask uint8 [M, N])."""
    m, n, k = xp.rows, wp.rows, wp.cols
    yp = alloc(m, n, xp.data.device, ones_col)
    mask = dropout_keep_mask((m, n), p, seed, offset, offset_dev, xp.data.device)   # the GEMM epilogue ANDs in (z > 0)
    epi = _dropout_epilogue(L.EPI_RELU_DROPOUT_PLANES, yp, mask, p, ones_col)
    gemm(xp, wp, m, n, k, bias=bias, split_k=1, epilogue=epi, want_out=False)
    return yp, mask


def relu_dropout_planes(z: torch.Tensor, p: float, seed: int, offset: int, offset_dev: Optional[torch.Tensor] = None,
                        ones_col: bool = True):
    """dropout_p(relu(z)) of an fp32 activation as (planes (+ ones column), keep-and-positive mask) in one HBM pass."""
    lib = L.load()
    m, n = z.shape
    yp = alloc(m, n, z.device, ones_col)
    mask = torch.empty(m, n, dtype=torch.uint8, device=z.device)
    RF._call("relu_dropout_planes", lib.rsb_relu_dropout_planes, L.ptr(z), m, n, z.stride(0), float(p),
             seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, L.ptr(offset_dev), int(ones_col), L.ptr(yp.data), yp.ld,
             yp.data.stride(0), L.ptr(mask), L.stream_ptr(z.device), nbytes=m * n * 5 + 3 * m * yp.ld * 2)
    return yp, mask


def dx_masked(gp: Planes, wtp: Planes, mask: torch.Tensor, p: float, to_planes: bool = True):
    """(g W) * mask / (1 - p) with `wtp` = planes of W^T [in, out] (K-major B): the gradient w.r.t. the previous layer's
    pre-activation, as planes (next GEMM operand) or as fp32."""
    m, n, k = gp.rows, wtp.rows, wtp.cols
    if to_planes:
        out = alloc(m, n, gp.data.device)
        epi = _dropout_epilogue(L.EPI_MASK_PLANES, out, mask, p, False)
        gemm(gp, wtp, m, n, k, split_k=1, epilogue=epi, want_out=False)
        return out
    epi = _dropout_epilogue(L.EPI_MASK_F32, None, mask, p, False)
    return gemm(gp, wtp, m, n, k, split_k=1, epilogue=epi)


def gemm_dw(gp: Planes, xp: Planes, want_bias_grad: bool):
    """Weight gradient g^T x (reduction over the batch rows of both operands, split over the SMs).  When `xp` carries a
    ones column (split(..., ones_col=True) / the fused forward epilogue) and want_bias_grad is set, the bias gradient
    sum_r g[r, :] falls out as one more output column: returns (dW [out, in], db [out] or None)."""
    n_out, n_in, m = gp.cols, xp.cols, gp.rows
    n_pad = (n_in + 7) // 8 * 8
    has_ones = xp.ld >= n_pad + 8
    if want_bias_grad and has_ones:
        full = gemm(gp, xp, n_out, n_pad + 8, m, a_mn_major=True, b_mn_major=True, split_k=0, b_cols=n_pad + 8)
        return full[:, :n_in], full[:, n_pad]
    return gemm(gp, xp, n_out, n_in, m, a_mn_major=True, b_mn_major=True, split_k=0), None


def rank1_mask_planes(g_row: torch.Tensor, w_col: torch.Tensor, mask: torch.Tensor, p: float) -> Planes:
    """Planes of gz[r, c] = g_row[r] * w_col[c] * mask[r, c] / (1 - p)."""
    lib = L.load()
    m, n = mask.shape
    out = alloc(m, n, mask.device)
    RF._call("rank1_mask_planes", lib.rsb_rank1_mask_planes, L.ptr(g_row), L.ptr(w_col), L.ptr(mask), m, n, float(p),
             L.ptr(out.data), out.ld, out.data.stride(0), L.stream_ptr(mask.device), nbytes=m * n * 7 + m * 4)
    return out


# ------------------------------------------------------------------ BatchNorm1d (training mode) around the GEMMs ---
def bn_train_stats(z: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float,
                   momentum: float, running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor]):
    """Batch statistics of z [M, N] -> (stats [2N] = mean | rstd, affine [2N] = scale | shift); running buffers are
    updated in place with torch.nn.BatchNorm1d's rule."""
    lib = L.load()
    m, n = z.shape
    dev = z.device
    stats = torch.empty(2 * n, dtype=torch.float32, device=dev)
    affine = torch.empty(2 * n, dtype=torch.float32, device=dev)
    ws = RF._ws(lib.rsb_bn_workspace_bytes(m, n), dev)
    RF._call("bn_fwd_stats", lib.rsb_bn_train_fwd_stats, L.ptr(z), m, n, z.stride(0), L.ptr(gamma), L.ptr(beta), float(eps),
             float(momentum), L.ptr(running_mean), L.ptr(running_var), L.ptr(stats), L.ptr(affine), L.ptr(ws), ws.numel(),
             L.stream_ptr(dev), nbytes=m * n * 4)
    return stats, affine


def bn_relu_dropout_planes(z: torch.Tensor, affine: torch.Tensor, p: float, seed: int, offset: int,
                           offset_dev: Optional[torch.Tensor] = None, ones_col: bool = True):
    """dropout_p(relu(z * scale + shift)) as (planes (+ ones column), keep-and-positive mask)."""
    lib = L.load()
    m, n = z.shape
    yp = alloc(m, n, z.device, ones_col)
    mask = torch.empty(m, n, dtype=torch.uint8, device=z.device)
    RF._call("bn_relu_dropout_planes", lib.rsb_bn_relu_dropout_planes, L.ptr(z), m, n, z.stride(0), L.ptr(affine), float(p),
             seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, L.ptr(offset_dev), int(ones_col), L.ptr(yp.data), yp.ld,
             yp.data.stride(0), L.ptr(mask), L.stream_ptr(z.device), nbytes=m * n * 5 + 3 * m * yp.ld * 2)
    return yp, mask


def bn_train_bwd_planes(g: torch.Tensor, z: torch.Tensor, stats: torch.Tensor, gamma: Optional[torch.Tensor]):
    """g = gradient w.r.t. the BatchNorm output -> (planes of the gradient w.r.t. z, d beta [N], d gamma [N])."""
    lib = L.load()
    m, n = z.shape
    dev = z.device
    sums = torch.empty(2 * n, dtype=torch.float32, device=dev)
    out = alloc(m, n, dev)
    ws = RF._ws(lib.rsb_bn_workspace_bytes(m, n), dev)
    RF._call("bn_bwd_planes", lib.rsb_bn_train_bwd_planes, L.ptr(g), L.ptr(z), m, n, g.stride(0), z.stride(0), L.ptr(stats),
             L.ptr(gamma), L.ptr(sums), L.ptr(out.data), out.ld, out.data.stride(0), L.ptr(ws), ws.numel(), L.stream_ptr(dev),
             nbytes=m * n * 16 + 3 * m * out.ld * 2)
    return out, sums[:n], sums[n:]
