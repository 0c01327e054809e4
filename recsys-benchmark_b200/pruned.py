"""Inference-only embedding over a pruned table stored as CSR.

Mirrors the reference's `PrunedEmbedding` (src/models/embeddings/pruned_embedding.py:11-138): same
constructors (`from_other_emb`, `from_weight`), `to_cuda()`, `get_weight()`, `forward(x)` for ids of any
shape, attributes `values` / `crow_indices` / `col_indices` / `is_cuda`.  The reference keeps numpy /
numba device arrays and launches a one-thread-per-id numba kernel; here the three arrays are torch
buffers (non-persistent: the reference's state dict has no entries for them either) and every
forward is one launch of `rsb_csr_lookup_fwd`; `lookup(x, offsets, fc, bias)` additionally fuses the
offsets add, DeepFM's first-order term and the FM second order, so an eval-mode DeepFM / DCN_Mix with
a pruned table is gather + MLP only (scripts/deepfm/infer_deepfm.py:138-153,307-352).

`compact=True` stores int32 row extents and uint8 columns (5 bytes per kept weight instead of 12).
"""
from __future__ import annotations

from typing import List, Optional, Union

import torch

from . import functional as RF
from .embeddings import IEmbedding


def _canonical_csr(weight: torch.Tensor) -> torch.Tensor:
    """CSR with sorted, unique columns per row (what the kernel's popcount expansion relies on)."""
    if weight.layout != torch.sparse_csr:
        return weight.to_sparse_csr()            # dense / COO -> CSR: row-major order, sorted columns
    col, crow = weight.col_indices(), weight.crow_indices()
    if col.numel() > 1:
        inc = col[1:] > col[:-1]
        starts = torch.zeros(col.numel() + 1, dtype=torch.bool, device=col.device)
        starts[crow.long().clamp(max=col.numel())] = True   # positions where a new row begins
        if not bool((inc | starts[1:-1]).all()):
            counts = (crow[1:] - crow[:-1]).long()
            rows = torch.repeat_interleave(torch.arange(counts.numel(), device=col.device), counts)
            order = torch.argsort(rows * weight.shape[1] + col.long(), stable=True)
            return torch.sparse_csr_tensor(crow, col[order], weight.values()[order], size=weight.shape)
    return weight


class PrunedEmbedding(IEmbedding):
    def __init__(self, field_dims: Union[int, List[int]], hidden_size: int, mode: Optional[str] = None):
        super().__init__()
        if isinstance(field_dims, int):
            field_dims = [field_dims]
        self._hidden_size = hidden_size
        self._num_item = sum(field_dims)
        self._mode = mode                       # kept for API parity; the reference's forward ignores it too
        self.register_buffer("values", torch.empty(0), persistent=False)
        self.register_buffer("crow_indices", torch.zeros(self._num_item + 1, dtype=torch.int64), persistent=False)
        self.register_buffer("col_indices", torch.empty(0, dtype=torch.int64), persistent=False)
        self.register_buffer("_rsb_err_flag", torch.zeros(1, dtype=torch.int32), persistent=False)

    @property
    def is_cuda(self) -> bool:
        return self.values.is_cuda

    @classmethod
    @torch.no_grad()
    def from_other_emb(cls, emb, mode=None, compact: bool = False) -> "PrunedEmbedding":
        return cls.from_weight(emb.get_weight(), mode, compact=compact)

    @classmethod
    @torch.no_grad()
    def from_weight(cls, weight: torch.Tensor, mode=None, compact: bool = False) -> "PrunedEmbedding":
        num_item, hidden_size = weight.shape
        result = cls(num_item, hidden_size, mode)
        csr = _canonical_csr(weight.detach())
        values = csr.values().to(torch.float32).contiguous()
        crow, col = csr.crow_indices().contiguous(), csr.col_indices().contiguous()
        if compact:
            if values.numel() >= 2 ** 31 or hidden_size > 256:
                raise ValueError("compact CSR needs nnz < 2^31 and hidden_size <= 256")
            crow, col = crow.to(torch.int32), col.to(torch.uint8)
        result.values, result.crow_indices, result.col_indices = values, crow, col
        result._rsb_err_flag = result._rsb_err_flag.to(values.device)
        return result

    def to_cuda(self):
        """Reference API (pruned_embedding.py:51-65)."""
        if not self.is_cuda:
            self.to("cuda")

    def get_num_params(self) -> int:
        return int(self.values.numel())

    def get_weight(self) -> torch.Tensor:
        csr = torch.sparse_csr_tensor(self.crow_indices.long(), self.col_indices.long(), self.values,
                                      size=(self._num_item, self._hidden_size), dtype=torch.float32)
        return csr.to_dense()

    def lookup(self, x: torch.Tensor, offsets: Optional[torch.Tensor] = None, fc: Optional[torch.Tensor] = None,
               bias: Optional[torch.Tensor] = None):
        if offsets is not None:
            offsets = offsets.reshape(-1)
            if offsets.dtype != torch.int64:
                offsets = offsets.long()
        emb, y = RF.csr_lookup(x, offsets, self.values, self.crow_indices, self.col_indices, self._num_item,
                               self._hidden_size, None if fc is None else fc.detach(),
                               None if bias is None else bias.detach(), self._rsb_err_flag)
        if self.validate:
            RF.check_index_errors(self)
        return emb, y

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """ids of any shape, already offset -> [*x.shape, hidden_size] (pruned_embedding.py:89-138)."""
        shape = x.shape
        x2 = x.reshape(-1, 1) if x.dim() <= 1 else x.reshape(-1, shape[-1])
        emb, _ = self.lookup(x2)
        return emb.reshape(*shape, self._hidden_size)
