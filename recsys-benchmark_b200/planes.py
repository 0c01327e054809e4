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

    def operand(self, mn_major: bool, row_step: int = 0, col_step: int = 0) -> L.PlanesOperand:
        return L.PlanesOperand(self.data.data_ptr(), self.rows, self.cols, self.ld, self.data.stride(0), int(mn_major),
                               row_step, col_step)

    def float(self) -> torch.Tensor:
        """The represented fp32 matrix (sum of the planes) - for tests."""
        return self.data.float().sum(0)[:, :self.cols]


def split(x: torch.Tensor, transpose: bool = False) -> Planes:
    """fp32 [rows, cols] (unit column stride) -> Planes of x (or of x^T)."""
    lib = L.load()
    dev = L.require_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    orows, ocols = (cols, rows) if transpose else (rows, cols)
    ld = (ocols + 7) // 8 * 8
    out = torch.empty(3, orows, ld, dtype=torch.bfloat16, device=dev)
    RF._call("split_planes", lib.rsb_split_planes, L.ptr(x), rows, cols, x.stride(0), int(transpose), L.ptr(out), ld,
             out.stride(0), L.stream_ptr(dev), nbytes=rows * cols * 4 + 3 * orows * ld * 2)
    return Planes(out, orows, ocols)


def gemm(a: Planes, b: Planes, m: int, n: int, k: int, *, a_mn_major: bool = False, b_mn_major: bool = False,
         bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 0.0,
         c: Optional[torch.Tensor] = None, split_k: int = 0, batch: int = 1, a_steps=(0, 0), b_steps=(0, 0),
         d_batch_stride: int = 0) -> torch.Tensor:
    """D [m, n] = alpha * A B^T + beta * C + bias on the tensor cores.

    a_mn_major = False: `a` stores [m, k];  True: `a` stores [k, m] (reduction over the stored rows).
    b_mn_major = False: `b` stores [n, k] (nn.Linear weight layout);  True: `b` stores [k, n].
    batch > 1: operand l starts at stored (row, col) + l * steps; D[l] = out.data + l * d_batch_stride floats."""
    lib = L.load()
    dev = a.data.device
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=dev)
    nb = lib.rsb_gemm_planes_workspace_bytes(m, n, k, batch, split_k)
    ws = RF._ws(nb, dev)
    oa = a.operand(a_mn_major, *a_steps)
    ob = b.operand(b_mn_major, *b_steps)
    RF._call("gemm_planes", lib.rsb_gemm_planes, C.byref(oa), C.byref(ob), m, n, k, batch, split_k, L.ptr(c), L.ptr(out),
             out.stride(0), d_batch_stride, L.ptr(bias), alpha, beta, L.ptr(ws), ws.numel(), L.stream_ptr(dev),
             nbytes=0)
    return out
