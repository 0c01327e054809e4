"""fp32 matrices held as 16-bit planes, and the hand-written tensor-core GEMM on them (rsb_gemm_planes).

`Planes` is the operand format of csrc/gemm/planes_gemm.cu, written ONCE per operand (by `split`, or by a producer
kernel's epilogue) and read by every GEMM that uses the operand - the forward GEMM and the weight-gradient GEMM of a
layer share the activation planes, the dX and dW GEMMs share the gradient planes, and the planes of an nn.Linear weight
are rebuilt only when the weight has changed.  Two formats (include/rsb.h, rsb_planes_format):

  BF16X3  X = X0 + X1 + X2, 8 mantissa bits per plane: every fp32 value exactly, 6 MMAs per product.
  FP16X2  X * s = H0 + H1, 11 bits per plane, s a power of two derived ON THE DEVICE from a bound on |X| (`amax`, a
          device scalar filled by rsb_absmax or by the BatchNorm statistics kernels - no host round trip): 22 bits
          for the large elements, 2^-38 of the bound for all, 3 MMAs per product."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from . import functional as RF


BF16X3, FP16X2 = L.PLANES_BF16X3, L.PLANES_FP16X2
MAX_SCALE_EXP = 40            # FP16X2 scale <= 2^40 (gradients of ~1e-8 still land in [2^13, 2^14))
ONES_SCALE_EXP = 14           # planes with a ones column store the scale itself in that column: it must fit fp16


class Planes:
    """16-bit [P, rows, ld] (ld = cols rounded up to 8; P = 3 bf16 or 2 fp16 planes) + the logical shape of the fp32
    matrix it represents; FP16X2 also carries the device scalar `amax` its scale derives from."""

    __slots__ = ("data", "rows", "cols", "fmt", "amax", "max_exp", "_ops")

    def __init__(self, data: torch.Tensor, rows: int, cols: int, fmt: int = BF16X3, amax: Optional[torch.Tensor] = None,
                 max_exp: int = MAX_SCALE_EXP):
        self.data, self.rows, self.cols, self.fmt, self.amax, self.max_exp = data, rows, cols, fmt, amax, max_exp
        self._ops = None      # operand descriptors already built for this (immutable) set of planes

    def format(self) -> L.PlanesFormat:
        return L.PlanesFormat(self.fmt, self.max_exp, self.amax.data_ptr() if self.amax is not None else None)

    def scale(self) -> torch.Tensor:
        """The FP16X2 scale as the kernels derive it (device scalar; tests and the bias-gradient fix-up)."""
        if self.fmt != FP16X2:
            return torch.ones((), device=self.data.device)
        a = self.amax.reshape(()).double()
        e = torch.floor(torch.log2(a)) + 1            # amax in [2^(e-1), 2^e)
        se = torch.clamp(14 - e, max=self.max_exp)
        return torch.where(a > 0, torch.exp2(se), torch.ones_like(a)).float()

    @property
    def ld(self) -> int:
        return self.data.shape[2]

    def operand(self, mn_major: bool, row_step: int = 0, col_step: int = 0, cols: Optional[int] = None) -> L.PlanesOperand:
        """`cols`: extent of the stored columns the GEMM may read (TMA zero-fills beyond it).  Default: the logical
        width - a forward GEMM must not reduce over the padding / ones column; a weight-gradient GEMM that wants the
        bias gradient passes the padded width + 8."""
        key = (mn_major, row_step, col_step, cols)
        ops = self._ops
        if ops is None:
            ops = self._ops = {}
        op = ops.get(key)
        if op is None:
            op = ops[key] = L.PlanesOperand(self.data.data_ptr(), self.rows, self.cols if cols is None else cols, self.ld,
                                            self.data.stride(0), int(mn_major), row_step, col_step, self.format())
        return op

    def float(self) -> torch.Tensor:
        """The represented fp32 matrix (sum of the planes, unscaled) - for tests."""
        x = self.data.double().sum(0)[:, :self.cols]
        return (x / self.scale().double()).float() if self.fmt == FP16X2 else x.float()


def alloc(rows: int, cols: int, device, ones_col: bool = False, fmt: int = BF16X3,
          amax: Optional[torch.Tensor] = None) -> Planes:
    """Uninitialised planes for a [rows, cols] matrix (the writer fills every column up to ld).  FP16X2: `amax` is the
    device scalar the writer's bound goes to / comes from (a fresh zero if not given)."""
    ld = (cols + 7) // 8 * 8 + (8 if ones_col else 0)
    if fmt == FP16X2:
        if amax is None:
            amax = torch.zeros(1, dtype=torch.float32, device=device)
        return Planes(torch.empty(2, rows, ld, dtype=torch.float16, device=device), rows, cols, FP16X2, amax,
                      ONES_SCALE_EXP if ones_col else MAX_SCALE_EXP)
    return Planes(torch.empty(3, rows, ld, dtype=torch.bfloat16, device=device), rows, cols)


def absmax(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """max |x| of an fp32 matrix as a device scalar [1] (raised into `out` if given)."""
    lib = L.load()
    dev = L.require_cuda(x)
    if out is None:
        out = torch.zeros(1, dtype=torch.float32, device=dev)
    RF._call("absmax", lib.rsb_absmax, L.ptr(x), x.shape[0], x.shape[1], x.stride(0), L.ptr(out), L.stream_ptr(dev),
             nbytes=x.numel() * 4)
    return out


def split(x: torch.Tensor, transpose: bool = False, ones_col: bool = False, fmt: int = BF16X3,
          amax: Optional[torch.Tensor] = None) -> Planes:
    """fp32 [rows, cols] (unit column stride) -> Planes of x (or of x^T).  ones_col: a column of ones after the
    8-padded data columns (see `gemm_dw`).  FP16X2: `amax` = an already known bound on |x| (device scalar), else one
    rsb_absmax pass computes it."""
    lib = L.load()
    dev = L.require_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    orows, ocols = (cols, rows) if transpose else (rows, cols)
    if fmt == FP16X2 and amax is None:
        amax = absmax(x)
    out = alloc(orows, ocols, dev, ones_col, fmt, amax)
    pf = out.format()
    RF._call("split_planes", lib.rsb_split_planes, L.ptr(x), rows, cols, x.stride(0), int(transpose), int(ones_col),
             L.ptr(out.data), out.ld, out.data.stride(0), C.byref(pf), L.stream_ptr(dev),
             nbytes=rows * cols * 4 + out.data.shape[0] * orows * out.ld * 2)
    return out


def _dropout_epilogue(mode: int, out: Optional[Planes], mask: torch.Tensor, p: float, ones_col: bool,
                      d_amax: Optional[torch.Tensor] = None) -> L.GemmEpilogue:
    return L.GemmEpilogue(mode, out.data.data_ptr() if out is not None else None, out.ld if out is not None else 0,
                          out.data.stride(0) if out is not None else 0, int(ones_col), mask.data_ptr(), float(p),
                          d_amax.data_ptr() if d_amax is not None else None)


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
            ("gemm_planes", "gemm_planes_relu_dropout", "gemm_planes_masked", "gemm_planes_masked")[epilogue.mode])
    if a_mn_major and b_mn_major:
        kind = "gemm_planes_dw"
    if a.fmt == FP16X2:
        kind += "_h"                               # 3 MMAs per product instead of 6 (bench.py's tensor roofline)
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


def dx_masked(gp: Planes, wtp: Planes, mask: torch.Tensor, p: float, to_planes: bool = True,
              d_amax: Optional[torch.Tensor] = None):
    """(g W) * mask / (1 - p) with `wtp` = planes of W^T [in, out] (K-major B): the gradient w.r.t. the previous layer's
    pre-activation, as planes (next GEMM operand) or as fp32.  d_amax (fp32 output only; zeroed device scalar): raised to
    max |result| by the GEMM epilogue - the bound the FP16X2 `bn_train_bwd_planes` of the previous layer needs."""
    m, n, k = gp.rows, wtp.rows, wtp.cols
    if to_planes:
        out = alloc(m, n, gp.data.device)
        epi = _dropout_epilogue(L.EPI_MASK_PLANES, out, mask, p, False)
        gemm(gp, wtp, m, n, k, split_k=1, epilogue=epi, want_out=False)
        return out
    epi = _dropout_epilogue(L.EPI_MASK_F32, None, mask, p, False, d_amax)
    return gemm(gp, wtp, m, n, k, split_k=1, epilogue=epi)


def rank1_absmax(g_row: torch.Tensor, w_col: torch.Tensor, p: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device scalar >= max |g_row[r] * w_col[c] * mask / (1 - p)| (the head gradient of rsb_relu_dropout_bwd_rank1)."""
    lib = L.load()
    dev = g_row.device
    if out is None:
        out = torch.zeros(1, dtype=torch.float32, device=dev)
    RF._call("rank1_absmax", lib.rsb_rank1_absmax, L.ptr(g_row), g_row.numel(), L.ptr(w_col), w_col.numel(),
             1.0 / (1.0 - p), L.ptr(out), L.stream_ptr(dev), nbytes=g_row.numel() * 4)
    return out


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
                   momentum: float, running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
                   act_amax: Optional[torch.Tensor] = None, bound_mul: float = 1.0):
    """Batch statistics of z [M, N] -> (stats [2N] = mean | rstd, affine [2N] = scale | shift); running buffers are
    updated in place with torch.nn.BatchNorm1d's rule.  act_amax (zeroed device scalar): receives bound_mul * max of
    relu(z * scale + shift), the bound of the FP16X2 planes `bn_relu_dropout_planes` writes next."""
    lib = L.load()
    m, n = z.shape
    dev = z.device
    stats = torch.empty(2 * n, dtype=torch.float32, device=dev)
    affine = torch.empty(2 * n, dtype=torch.float32, device=dev)
    ws = RF._ws(lib.rsb_bn_workspace_bytes(m, n), dev)
    RF._call("bn_fwd_stats", lib.rsb_bn_train_fwd_stats, L.ptr(z), m, n, z.stride(0), L.ptr(gamma), L.ptr(beta), float(eps),
             float(momentum), L.ptr(running_mean), L.ptr(running_var), L.ptr(stats), L.ptr(affine), float(bound_mul),
             L.ptr(act_amax), L.ptr(ws), ws.numel(), L.stream_ptr(dev), nbytes=m * n * 4)
    return stats, affine


def gemm_bn_stats(xp: Planes, wp: Planes, bias: Optional[torch.Tensor], gamma: Optional[torch.Tensor],
                  beta: Optional[torch.Tensor], eps: float, momentum: float, running_mean: Optional[torch.Tensor],
                  running_var: Optional[torch.Tensor], act_amax: Optional[torch.Tensor] = None, bound_mul: float = 1.0,
                  in_epilogue: bool = True):
    """z = x W^T + b AND the BatchNorm batch statistics of z: -> (z, stats, affine) like `gemm` + `bn_train_stats`.
    in_epilogue (and the GEMM can carry them: fp32 TMA-store epilogue): the statistics are reduced from the staged
    output tiles in the GEMM's epilogue (per-32-row-group shifted sums) and only a small combination follows - no pass
    over z.  Otherwise: the GEMM, then rsb_bn_train_fwd_stats."""
    lib = L.load()
    m, n, k = xp.rows, wp.rows, wp.cols
    dev = xp.data.device
    nb = lib.rsb_gemm_bn_partials_bytes(m, n, k, xp.fmt) if in_epilogue else -1
    if nb <= 0:
        z = gemm(xp, wp, m, n, k, bias=bias, split_k=1)
        return (z, *bn_train_stats(z, gamma, beta, eps, momentum, running_mean, running_var, act_amax, bound_mul))
    parts = torch.empty(nb // 4, dtype=torch.float32, device=dev)
    epi = L.GemmEpilogue(L.EPI_LINEAR, None, 0, 0, 0, None, 0.0, None, parts.data_ptr())
    z = gemm(xp, wp, m, n, k, bias=bias, split_k=1, epilogue=epi)
    stats = torch.empty(2 * n, dtype=torch.float32, device=dev)
    affine = torch.empty(2 * n, dtype=torch.float32, device=dev)
    ws = RF._ws(16 * 2 * n * 4 + 512, dev)
    RF._call("bn_finalize_partials", lib.rsb_bn_finalize_partials, L.ptr(parts), m, n, L.ptr(gamma), L.ptr(beta), float(eps),
             float(momentum), L.ptr(running_mean), L.ptr(running_var), L.ptr(stats), L.ptr(affine), float(bound_mul),
             L.ptr(act_amax), L.ptr(ws), ws.numel(), L.stream_ptr(dev), nbytes=parts.numel() * 4)
    return z, stats, affine


def bn_relu_dropout_planes(z: torch.Tensor, affine: torch.Tensor, p: float, seed: int, offset: int,
                           offset_dev: Optional[torch.Tensor] = None, ones_col: bool = True, fmt: int = BF16X3,
                           amax: Optional[torch.Tensor] = None):
    """dropout_p(relu(z * scale + shift)) as (planes (+ ones column), keep-and-positive mask).  FP16X2: `amax` is the
    bound `bn_train_stats(..., act_amax=amax, bound_mul=1 / (1 - p))` produced."""
    lib = L.load()
    m, n = z.shape
    assert fmt == BF16X3 or amax is not None
    yp = alloc(m, n, z.device, ones_col, fmt, amax)
    pf = yp.format()
    mask = torch.empty(m, n, dtype=torch.uint8, device=z.device)
    RF._call("bn_relu_dropout_planes", lib.rsb_bn_relu_dropout_planes, L.ptr(z), m, n, z.stride(0), L.ptr(affine), float(p),
             seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, L.ptr(offset_dev), int(ones_col), L.ptr(yp.data), yp.ld,
             yp.data.stride(0), L.ptr(mask), C.byref(pf), L.stream_ptr(z.device),
             nbytes=m * n * 5 + yp.data.shape[0] * m * yp.ld * 2)
    return yp, mask


def bn_train_bwd_planes(g: torch.Tensor, z: torch.Tensor, stats: torch.Tensor, gamma: Optional[torch.Tensor],
                        fmt: int = BF16X3, g_amax: Optional[torch.Tensor] = None):
    """g = gradient w.r.t. the BatchNorm output -> (planes of the gradient w.r.t. z, d beta [N], d gamma [N]).
    FP16X2: `g_amax` = device scalar >= max |g| (from g's producer: `dx_masked(..., d_amax=)` / `rank1_absmax`; computed
    by one rsb_absmax pass if not given); the scale of gz comes from |gz| <= |gamma rstd| * g_amax * (2 + sqrt(M))."""
    lib = L.load()
    m, n = z.shape
    dev = z.device
    sums = torch.empty(2 * n, dtype=torch.float32, device=dev)
    if fmt == FP16X2 and g_amax is None:
        g_amax = absmax(g)
    out = alloc(m, n, dev, fmt=fmt)
    pf = out.format()
    ws = RF._ws(lib.rsb_bn_workspace_bytes(m, n), dev)
    RF._call("bn_bwd_planes", lib.rsb_bn_train_bwd_planes, L.ptr(g), L.ptr(z), m, n, g.stride(0), z.stride(0), L.ptr(stats),
             L.ptr(gamma), L.ptr(sums), L.ptr(out.data), out.ld, out.data.stride(0), C.byref(pf), L.ptr(g_amax), L.ptr(ws),
             ws.numel(), L.stream_ptr(dev), nbytes=m * n * 16 + out.data.shape[0] * m * out.ld * 2)
    return out, sums[:n], sums[n:]
