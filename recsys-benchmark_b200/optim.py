"""Fused sparse row optimizers: the sorted segmented reduction applies the update in the
same pass that sums the duplicate rows, so no gradient tensor (dense or COO) is ever
materialised for the table.

Arithmetic parity: torch.optim.SparseAdam (torch/optim/_functional.py:24-84) and sparse
torch.optim.SGD as selected by the reference at src/models/deepfm.py:173-216 — touched
rows only, global step for the bias correction, no weight decay.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib as L
from . import functional as RF


class _FusedRowOptimizer(torch.optim.Optimizer):
    """Owns the embedding module's main table.  While attached, the module's backward
    stashes (rows, per-lookup grads) here instead of building a gradient tensor.

    The table's `.grad` therefore stays None: `clip_grad_norm_(model.parameters())` neither counts nor clips the
    embedding gradient in this mode.  The reference cannot clip there either (clip_grad_norm_ raises on the COO
    gradient of `sparse=True`, which is why its sparse configs run with clip_grad = 0, src/trainer/deepfm.py:24,56)."""

    def __init__(self, embedding_module, defaults):
        table, table1, _aux = embedding_module._tensors()
        if table1 is not None or embedding_module._spec().kind not in (L.KIND_VANILLA, L.KIND_MASK):
            raise ValueError("fused sparse updates support vanilla / masked single-table embeddings")
        super().__init__([table], defaults)
        self._module = embedding_module
        self._pending: List[Tuple] = []
        embedding_module._rsb_fused_opt = self

    def stash(self, table, rows, row_grads, sorted_pair=None):
        self._pending.append((table, rows, row_grads, sorted_pair))

    def _take_pending(self):
        """All backward passes since the last step as ONE gradient (what torch.optim.SparseAdam / SGD see when
        several backwards accumulate into one COO .grad: it coalesces and steps once).  A single backward keeps its
        presorted rows; several are concatenated and sorted together."""
        pending, self._pending = self._pending, []
        if len(pending) <= 1:
            return pending
        table = pending[0][0]
        rows = torch.cat([p[1].reshape(-1) for p in pending])
        rg = torch.cat([p[2].reshape(-1, p[2].shape[-1]) for p in pending])
        return [(table, rows, rg, None)]

    def zero_grad(self, set_to_none: bool = True):
        self._pending.clear()
        super().zero_grad(set_to_none)

    def detach_from_module(self):
        if getattr(self._module, "_rsb_fused_opt", None) is self:
            self._module._rsb_fused_opt = None


class FusedSparseAdam(_FusedRowOptimizer):
    def __init__(self, embedding_module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(embedding_module, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        group = self.param_groups[0]
        p = group["params"][0]
        for table, rows, rg, pair in self._take_pending():
            state = self.state[p]
            if len(state) == 0:
                state["step"] = 0
                state["exp_avg"] = torch.zeros_like(p)
                state["exp_avg_sq"] = torch.zeros_like(p)
            state["step"] += 1
            b1, b2 = group["betas"]
            skeys, perm = pair if pair is not None else RF.sort_rows(rows, p.shape[0])
            RF.segment_reduce_apply(L.APPLY_SPARSE_ADAM, skeys, perm, rg, p.data, state["exp_avg"],
                                    state["exp_avg_sq"], lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"],
                                    step=state["step"])
        return None


class FusedSparseSGD(_FusedRowOptimizer):
    def __init__(self, embedding_module, lr=1e-3):
        super().__init__(embedding_module, dict(lr=lr))

    @torch.no_grad()
    def step(self, closure=None):
        group = self.param_groups[0]
        p = group["params"][0]
        for table, rows, rg, pair in self._take_pending():
            skeys, perm = pair if pair is not None else RF.sort_rows(rows, p.shape[0])
            RF.segment_reduce_apply(L.APPLY_SPARSE_SGD, skeys, perm, rg, p.data, lr=group["lr"])
        return None
