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


class FusedDenseAdam(torch.optim.Optimizer):
    """Dense Adam over all its parameters in ONE launch of `rsb_adam_dense` (csrc/staging.cu).

    Arithmetic parity: `torch.optim.Adam(params, lr, betas, eps, weight_decay)` as the reference's `get_optimizers`
    builds it for the non-sparse configs (src/models/deepfm.py:155-172): L2-style weight decay added to the gradient,
    a step count per parameter, bias corrections formed in double on the host.  Parameters whose `.grad` is None are
    skipped like torch does (e.g. DeepFM.linear_layer, which never receives a gradient); parameters with equal step
    counts share one launch.  Selected with the opt-in config key `fused_adam: "rsb"`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._maps = {}       # numels -> block map on the device (depends on the sizes only)
        self._plans = {}      # group index -> launches for the current set of parameters that receive gradients

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans.clear()

    def _block_map(self, numels, dev):
        key = (numels, dev)
        bm = self._maps.get(key)
        if bm is None:
            parts = []
            for ti, n in enumerate(numels):
                nchunks = (n + L.ADAM_CHUNK - 1) // L.ADAM_CHUNK
                parts.append(torch.stack([torch.full((nchunks,), ti, dtype=torch.int32),
                                          torch.arange(nchunks, dtype=torch.int32)], 1))
            bm = self._maps[key] = torch.cat(parts).contiguous().to(dev)
        return bm

    def _plan(self, gi, group, mask):
        """Launches for the parameters of `group` selected by `mask`: torch counts steps per parameter, so parameters
        that skipped steps form their own launch; <= RSB_ADAM_MAX_TENSORS descriptors travel with one launch."""
        by_step = {}
        for p, has in zip(group["params"], mask):
            if not has:
                continue
            st = self.state[p]
            if len(st) == 0:
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedDenseAdam needs contiguous fp32 parameters")
                L.require_cuda(p)
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            by_step.setdefault((int(st["step"]), p.device), []).append(p)
        launches = []
        for key in sorted(by_step, key=lambda k: (k[0], str(k[1]))):
            step, plist = key[0], by_step[key]
            for lo in range(0, len(plist), L.ADAM_MAX_TENSORS):
                part = plist[lo:lo + L.ADAM_MAX_TENSORS]
                desc = (L.AdamTensor * len(part))()
                states = []
                for d, p in zip(desc, part):
                    st = self.state[p]
                    d.exp_avg, d.exp_avg_sq, d.numel = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
                    states.append(st)
                numels = tuple(p.numel() for p in part)
                launches.append(dict(step=step, params=part, states=states, desc=desc, n=len(part),
                                     bm=self._block_map(numels, part[0].device), nbytes=28 * sum(numels),
                                     device=part[0].device))
        plan = self._plans[gi] = dict(mask=mask, launches=launches)
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            mask = tuple(p.grad is not None for p in group["params"])
            plan = self._plans.get(gi)
            if plan is None or plan["mask"] != mask:
                plan = self._plan(gi, group, mask)
            b1, b2 = group["betas"]
            for ln in plan["launches"]:
                dev = ln["device"]
                for d, p in zip(ln["desc"], ln["params"]):
                    g = p.grad
                    if g.dtype is not torch.float32 or g.layout is not torch.strided or g.device != dev or not g.is_contiguous():
                        raise RuntimeError("FusedDenseAdam needs dense contiguous fp32 gradients on the parameter's device")
                    d.param = p.data_ptr()
                    d.grad = g.data_ptr()
                step = ln["step"] = ln["step"] + 1
                for st in ln["states"]:
                    st["step"] = step
                # p, g, m, v read + p, m, v written
                RF._call("adam_dense", lib.rsb_adam_dense, ln["desc"], ln["n"], L.ptr(ln["bm"]), int(ln["bm"].shape[0]),
                         float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), step,
                         L.stream_ptr(dev), nbytes=ln["nbytes"])
        return loss
