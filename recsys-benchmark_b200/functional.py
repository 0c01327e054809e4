"""Host-side glue between torch autograd and the C ABI (include/rsb.h).

`fused_lookup` is the one differentiable entry point: gather (+ lightweight-embedding
variant) (+ DeepFM first/second order) in one launch forward, and the three-stage
backward (per-lookup row grads -> row sort -> deterministic segmented reduction with
the consumer fused in).  torch is used for memory, streams and autograd plumbing only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib as L


@dataclass
class LookupSpec:
    """Static description of one embedding variant (what the gather applies)."""

    kind: int                      # L.KIND_*
    num_global: int                # number of addressable ids (sum(field_dims))
    dim: int                       # D (num_factor)
    divider: int = 0               # QR quotient divisor (CERP: q_entity_per_row)
    modulus: int = 0               # remainder modulus if different from divider (CERP bucket size)
    aux_mode: int = 0              # PEP threshold type / OptEmbed norm
    sparse_grad: bool = False      # emit torch.sparse_coo grads for the main table (nn.Embedding(sparse=True))
    module: object = None          # owner (for deferred fused updates / err flag)

    @property
    def is_qr(self) -> bool:
        return self.kind in (L.KIND_QR_MULT, L.KIND_QR_ADD, L.KIND_QR_CAT)

    @property
    def row_width(self) -> int:
        return self.dim // 2 if self.kind == L.KIND_QR_CAT else self.dim

    def out_fields(self, f: int) -> int:
        return 2 * f if self.kind == L.KIND_QR_CAT else f


class KernelTimer:
    """Optional per-C-call CUDA-event timing (bench.py's roofline leg).  Events are recorded
    on the launching stream right around each call; nothing synchronises until `summary()`."""

    def __init__(self):
        self.records = {}

    def add(self, name, start, end, nbytes=0):
        self.records.setdefault(name, []).append((start, end, nbytes))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = [s.elapsed_time(e) for s, e, _ in recs]
            out[name] = dict(calls=len(ms), ms_avg=sum(ms) / len(ms), ms_total=sum(ms),
                             bytes_avg=sum(r[2] for r in recs) / len(recs))
        return out


_TIMER = None


def set_timer(timer) -> None:
    global _TIMER
    _TIMER = timer


def _call(name: str, fn, *args, nbytes: int = 0) -> None:
    """One C-ABI call, optionally bracketed by CUDA events; raises on a non-zero status."""
    t = _TIMER
    if t is None:
        L.check(fn(*args), name)
        return
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn(*args)
    e.record()
    L.check(rc, name)
    t.add(name, s, e, nbytes)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------
# thin wrappers (one C call each)
# ----------------------------------------------------------------------------
def sort_rows(rows: torch.Tensor, n_rows: int, key_div: int = 0, key_mod: int = 0):
    """Stable radix sort of int64 row ids -> (sorted_keys uint32-as-int32, perm)."""
    lib = L.load()
    n = rows.numel()
    dev = rows.device
    skeys = torch.empty(n, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    nb = lib.rsb_sort_workspace_bytes(n)
    ws = _ws(nb, dev)
    _call("sort_rows", lib.rsb_sort_rows, L.ptr(rows), n, int(n_rows), int(key_div), int(key_mod), L.ptr(skeys),
          L.ptr(perm), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    return skeys, perm


# The gather kernels can report max |emb| (per-warp maxima in RSB_LOOKUP_AMAX_SLOTS slots): the dense tail's FP16X2
# operand split needs that bound, and getting it here saves a pass over the [B, F*D] activation.
EMIT_AMAX = True


def amax_slots_for(device) -> Optional[torch.Tensor]:
    """Zeroed slot array for a training-mode gather (None when off / not training)."""
    if not EMIT_AMAX or not torch.is_grad_enabled():
        return None
    return torch.zeros(L.LOOKUP_AMAX_SLOTS, dtype=torch.float32, device=device)


def amax_slots_of(emb: torch.Tensor) -> Optional[torch.Tensor]:
    return getattr(emb, "_rsb_amax_slots", None)


# ---- backward stage 2 started early --------------------------------------------------
# The sort needs only the looked-up row ids, which exist as soon as the forward gather has run, while its
# consumer (the segmented reduction) runs at the very end of the backward pass.  It is therefore queued on a
# side stream right after the forward kernel and overlaps the dense tail (GEMMs) of the forward/backward pass;
# the backward waits on its event.  Set EARLY_SORT = False to sort inside the backward instead.
EARLY_SORT = True
_SIDE_STREAMS = {}


def side_stream(device) -> torch.cuda.Stream:
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


class SortedRows:
    """(sorted keys, permutation) being produced on the side stream; `get()` makes the current stream wait."""

    constructed = 0   # instances ever made (tests assert that the early sort really runs)

    def __init__(self, rows: torch.Tensor, n_rows: int, key_div: int = 0, key_mod: int = 0):
        SortedRows.constructed += 1
        lib = L.load()
        dev = rows.device
        n = rows.numel()
        self.key = (int(n_rows), int(key_div), int(key_mod))
        # buffers come from the CURRENT stream's pool (they are consumed there); the side stream only borrows them
        self.skeys = torch.empty(n, dtype=torch.int32, device=dev)
        self.perm = torch.empty(n, dtype=torch.int32, device=dev)
        self._ws = _ws(lib.rsb_sort_workspace_bytes(n), dev)
        self._rows = rows
        main = torch.cuda.current_stream(dev)
        side = side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _call("sort_rows", lib.rsb_sort_rows, L.ptr(rows), n, int(n_rows), int(key_div), int(key_mod),
                  L.ptr(self.skeys), L.ptr(self.perm), L.ptr(self._ws), self._ws.numel(), L.stream_ptr(dev))
            self.event = torch.cuda.Event()
            self.event.record(side)
        self._joined = False
        # Lifetime instead of record_stream (which makes the caching allocator poll events and fall back to
        # cudaMalloc when the host runs ahead): this object keeps rows / workspace / outputs alive until the
        # consuming stream has been made to wait for the sort - in get(), or at destruction if never consumed.

    def get(self):
        torch.cuda.current_stream(self.skeys.device).wait_event(self.event)
        self._joined = True
        return self.skeys, self.perm

    def __del__(self):
        try:
            if not self._joined:
                torch.cuda.current_stream(self.skeys.device).wait_event(self.event)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def early_sort(rows: torch.Tensor, n_rows: int, key_div: int = 0, key_mod: int = 0) -> Optional[SortedRows]:
    if not EARLY_SORT or rows.numel() == 0:
        return None
    return SortedRows(rows, n_rows, key_div, key_mod)


def segment_reduce_apply(apply: int, skeys, perm, row_grads, dst, exp_avg=None, exp_avg_sq=None, lr=0.0,
                         beta1=0.9, beta2=0.999, eps=1e-8, step=1):
    lib = L.load()
    n = skeys.numel()
    e = row_grads.shape[-1] if row_grads.dim() > 1 else 1
    dev = row_grads.device
    nb = lib.rsb_segment_workspace_bytes(n, e)
    ws = _ws(nb, dev)
    _call("segment_reduce_apply", lib.rsb_segment_reduce_apply, apply, L.ptr(skeys), L.ptr(perm), n,
          L.ptr(row_grads), e, L.ptr(dst), L.ptr(exp_avg), L.ptr(exp_avg_sq), lr, beta1, beta2, eps, int(step),
          L.ptr(ws), ws.numel(), L.stream_ptr(dev), nbytes=n * (8 + 4 * e))


def dense_row_grad(rows: torch.Tensor, row_grads: torch.Tensor, n_rows: int, key_div: int = 0, key_mod: int = 0,
                   sorted_pair=None) -> torch.Tensor:
    """Dense zero-filled [n_rows, E] gradient = scatter-add of per-lookup rows (deterministic)."""
    e = row_grads.shape[-1]
    out = torch.zeros(n_rows, e, dtype=torch.float32, device=row_grads.device)
    if rows.numel() == 0:
        return out
    if sorted_pair is None:
        sorted_pair = sort_rows(rows, n_rows, key_div, key_mod)
    segment_reduce_apply(L.APPLY_DENSE, sorted_pair[0], sorted_pair[1], row_grads, out)
    return out


def small_table_grad(rows: torch.Tensor, row_grads: torch.Tensor, n_rows: int, key_div: int = 0,
                     key_mod: int = 0) -> Optional[torch.Tensor]:
    """Dense gradient of a table small enough to accumulate in shared memory; None if too big."""
    lib = L.load()
    e = row_grads.shape[-1]
    nb = lib.rsb_small_table_workspace_bytes(n_rows, e)
    if nb < 0:
        return None
    dev = row_grads.device
    out = torch.empty(n_rows, e, dtype=torch.float32, device=dev)
    ws = _ws(nb, dev)
    _call("small_table_grad", lib.rsb_small_table_grad, L.ptr(rows), rows.numel(), int(key_div), int(key_mod),
          L.ptr(row_grads), e, int(n_rows), L.ptr(out), L.ptr(ws), ws.numel(), L.stream_ptr(dev),
          nbytes=rows.numel() * (8 + 4 * e))
    return out


# The atomic first-order gradient kernel and the shared-memory atomics of the mid-size QR remainder table are the only
# order-dependent float sums of the backward pass.  The sorted first-order gradient is used whenever the full-row sort
# exists anyway (vanilla / PEP / OptEmbed / sharded tables); DETERMINISTIC = True also pays an extra sort where it
# does not (QR keys are sorted by quotient; COO gradients are not sorted at all) and skips the shared-memory path.
DETERMINISTIC = False


def fc_grad(rows: torch.Tensor, g_y: torch.Tensor, b: int, f: int, shape, n_rows: int, sorted_pair=None) -> torch.Tensor:
    """Dense gradient of the first-order weights fc [N,1] (src/models/deepfm.py:71-76): g_fc[row] = sum of g_y[b]
    over the lookups of `row`.  With the row-sorted lookups: one writer per row, fixed order (bit-reproducible)."""
    lib = L.load()
    dev = rows.device
    n = b * f
    g_fc = torch.zeros(shape, dtype=torch.float32, device=dev)
    if sorted_pair is None and DETERMINISTIC and n > 0:
        sorted_pair = sort_rows(rows, n_rows)
    if sorted_pair is not None:
        ws = _ws(lib.rsb_segment_workspace_bytes(n, 1), dev)
        _call("fc_grad_sorted", lib.rsb_fc_grad_sorted, L.ptr(sorted_pair[0]), L.ptr(sorted_pair[1]), n, L.ptr(g_y), f,
              L.ptr(g_fc), L.ptr(ws), ws.numel(), L.stream_ptr(dev), nbytes=n * 12)
    else:
        _call("fc_grad", lib.rsb_fc_grad, L.ptr(rows), L.ptr(g_y), b, f, L.ptr(g_fc), L.stream_ptr(dev), nbytes=n * 12)
    return g_fc


# The sorted first-order gradient (3 short, latency-bound launches) depends only on the sorted lookups and g_y, like
# the table's segmented reduction it used to follow: it runs BESIDE that reduction on its own stream and re-joins at
# the end of the backward pass.  The gradient is handed to autograd while still being written, so this is used only when
# AccumulateGrad will steal it (fc.grad is None, no hooks - the same rule as linalg._side_dw_safe), never under capture.
SIDE_FC_GRAD = True
_FC_STREAMS = {}


def fc_grad_beside(fc_param, rows, g_y, b, f, shape, n_rows, sorted_pair):
    safe = (SIDE_FC_GRAD and sorted_pair is not None and fc_param is not None and fc_param.grad is None
            and not getattr(fc_param, "_backward_hooks", None) and not getattr(fc_param, "_post_accumulate_grad_hooks", None)
            and not torch.cuda.is_current_stream_capturing() and _TIMER is None)
    if not safe:
        return fc_grad(rows, g_y, b, f, shape, n_rows, sorted_pair)
    dev = rows.device
    key = dev.index
    if key not in _FC_STREAMS:
        _FC_STREAMS[key] = torch.cuda.Stream(dev)
    side = _FC_STREAMS[key]
    main = torch.cuda.current_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        out = fc_grad(rows, g_y, b, f, shape, n_rows, sorted_pair)
    keep = (rows, g_y, sorted_pair)       # main-pool tensors the side stream reads: alive until the join
    try:
        torch.autograd.Variable._execution_engine.queue_callback(lambda keep=keep: main.wait_stream(side))
    except Exception:  # noqa: BLE001 - not inside an autograd pass: join right away
        main.wait_stream(side)
    return out


def _err_flag(spec: LookupSpec, device) -> Optional[torch.Tensor]:
    mod = spec.module
    if mod is None:
        return None
    flag = getattr(mod, "_rsb_err_flag", None)
    if flag is None or flag.device != device:
        flag = torch.zeros(1, dtype=torch.int32, device=device)
        mod._rsb_err_flag = flag
    return flag


def check_index_errors(module) -> None:
    """Raise IndexError (like F.embedding does) if any id seen so far was out of range.
    Synchronises the stream; called on demand (validate=True) or by tests."""
    flag = getattr(module, "_rsb_err_flag", None)
    if flag is not None and int(flag.item()) != 0:
        flag.zero_()
        raise IndexError("index out of range in self")


def _aux_row_bytes(spec: LookupSpec, f: int, mask_d) -> int:
    """Per-sample bytes of the per-row aux array a lookup reads beside the table rows: PEP thresholds of the
    feature_dim (a second row) / feature (4 B) kinds, the retrain mask (1 B per element), OptEmbed's mask-D ids."""
    if spec.kind == L.KIND_PEP:
        return f * spec.dim * 4 if spec.aux_mode == L.PEP_FEATURE_DIM else (f * 4 if spec.aux_mode == L.PEP_FEATURE else 0)
    if spec.kind == L.KIND_MASK:
        return f * spec.dim
    if spec.kind == L.KIND_OPTEMBED and mask_d is not None:
        return f * 8
    return 0


# ----------------------------------------------------------------------------
# the differentiable op
# ----------------------------------------------------------------------------
class _FusedLookup(torch.autograd.Function):
    """(spec, x, offsets, mask_d, table, table1, aux, fc, bias) -> (emb [B,VF,E], y_fm [B] or empty)."""

    @staticmethod
    def forward(ctx, spec: LookupSpec, x, offsets, mask_d, table, table1, aux, fc, bias, presort=False, amax_slots=None):
        lib = L.load()
        dev = L.require_cuda(x, table, table1, aux, fc, bias, offsets, mask_d)
        if x.dim() != 2:
            raise RuntimeError("rsb: x must be [B, F]")
        if x.dtype not in (torch.int32, torch.int64):
            raise RuntimeError(f"rsb: ids must be int32 or int64, got {x.dtype}")
        x = x.contiguous()
        b, f = x.shape
        e = spec.row_width
        vf = spec.out_fields(f)
        fm = fc is not None
        emb = torch.empty(b, vf, e, dtype=torch.float32, device=dev)
        y = torch.empty(b, dtype=torch.float32, device=dev) if fm else None
        s = torch.empty(b, e, dtype=torch.float32, device=dev) if fm else None
        rows = torch.empty(b, f, dtype=torch.int64, device=dev)
        aux_t = aux
        if aux is not None and aux.dtype == torch.bool:
            aux_t = aux.view(torch.uint8)
        aux_mode = spec.modulus if spec.is_qr else spec.aux_mode
        # algorithmic bytes: ids + rows read + emb written + rows saved (+ fc, y, S)
        r_bytes = vf * e * 4
        nbytes = b * (f * x.element_size() + 2 * r_bytes + _aux_row_bytes(spec, f, mask_d) + f * 8
                      + (f * 4 + 4 + e * 4 if fm else 0))
        _call("lookup_fwd", lib.rsb_lookup_fwd,
              spec.kind, L.ptr(x), int(x.dtype == torch.int32), L.ptr(offsets), b, f, spec.dim,
              L.ptr(table), table.shape[0], spec.num_global, L.ptr(table1), spec.divider,
              L.ptr(aux_t), aux_mode, L.ptr(mask_d), L.ptr(fc), L.ptr(bias),
              L.ptr(emb), L.ptr(y), L.ptr(s), L.ptr(rows), L.ptr(_err_flag(spec, dev)), L.ptr(amax_slots),
              L.stream_ptr(dev), nbytes=nbytes)
        ctx.spec = spec
        ctx.fm = fm
        ctx.fc_param = fc
        ctx.shape = (b, f)
        # the backward's row sort depends only on `rows`: start it now on the side stream (see EARLY_SORT)
        # (grad mode is always off inside Function.forward: the caller decides, see fused_lookup)
        ctx.presorted = None
        if presort:
            ctx.presorted = early_sort(rows, table.shape[0], key_div=spec.divider if spec.is_qr else 0)
        ctx.save_for_backward(rows, emb, s, mask_d, table, table1, aux, fc)
        ctx.mark_non_differentiable(rows)
        if y is None:
            y = emb.new_empty(0)
        return emb, y, rows

    @staticmethod
    def backward(ctx, g_emb, g_y, _g_rows):
        lib = L.load()
        spec: LookupSpec = ctx.spec
        rows, emb, s, mask_d, table, table1, aux, fc = ctx.saved_tensors
        b, f = ctx.shape
        dev = rows.device
        e = spec.row_width
        n = b * f
        need = ctx.needs_input_grad  # (spec, x, offsets, mask_d, table, table1, aux, fc, bias)
        use_gy = ctx.fm and g_y is not None and g_y.numel() == b
        if g_emb is not None:
            g_emb = g_emb.contiguous()
        if use_gy:
            g_y = g_y.contiguous()
        if g_emb is None and not use_gy:
            return (None,) * 11

        g_fc = None
        want_fc = ctx.fm and need[7] and use_gy
        g_bias = g_y.sum().reshape(1) if (ctx.fm and need[8] and use_gy) else None

        kind = spec.kind
        aux_t = aux.view(torch.uint8) if (aux is not None and aux.dtype == torch.bool) else aux
        # ---- stage 1: per-lookup row gradients --------------------------------
        skip_stage1 = (kind == L.KIND_VANILLA and not use_gy)
        rg_main = rg_aux = None
        g_table1_fused = None
        if skip_stage1:
            rg_main = g_emb.view(n, e)
        elif kind in (L.KIND_QR_MULT, L.KIND_QR_ADD) and spec.divider <= 8 and spec.modulus == 0 and need[5]:
            # emb1 (<= 8 rows) gradient accumulated in registers inside stage 1: no per-lookup emb1 rows
            rg_main = torch.empty(n, e, dtype=torch.float32, device=dev)
            g_table1_fused = torch.empty_like(table1)
            ws = _ws(lib.rsb_qr_bwd_fused_workspace_bytes(b, spec.dim), dev)
            r_bytes = f * e * 4
            nbytes = b * (f * 8 + r_bytes * (1 + int(use_gy)) + f * e * 4 + (e * 4 + 4 + f * 4 if use_gy else 0))
            _call("lookup_bwd_rows", lib.rsb_qr_bwd_fused, kind, L.ptr(rows), b, f, spec.dim, L.ptr(table),
                  table.shape[0], L.ptr(table1), spec.divider, L.ptr(emb), L.ptr(s), L.ptr(g_y) if use_gy else None,
                  L.ptr(g_emb), L.ptr(rg_main), L.ptr(g_table1_fused), None, L.ptr(ws), ws.numel(),
                  L.stream_ptr(dev), nbytes=nbytes)
        else:
            rg_main = torch.empty(n, e, dtype=torch.float32, device=dev)
            if kind in (L.KIND_QR_MULT, L.KIND_QR_CAT) or (kind == L.KIND_PEP and need[6]):
                rg_aux = torch.empty(n, e, dtype=torch.float32, device=dev)
            elif kind == L.KIND_OPTEMBED and aux is not None:
                rg_aux = torch.empty(b, f, dtype=torch.float32, device=dev)
            # algorithmic bytes: rows + g_deep + emb (+S, g_y) read, row grads written (+ fc)
            r_bytes = spec.out_fields(f) * e * 4
            n_out = 2 if (rg_aux is not None and kind != L.KIND_OPTEMBED) else 1
            # the PEP / OptEmbed chain rules re-read the weight rows (and the per-row thresholds / masks)
            reread = (r_bytes if kind in (L.KIND_PEP, L.KIND_OPTEMBED) else 0) + _aux_row_bytes(spec, f, mask_d)
            nbytes = b * (f * 8 + r_bytes * (1 + int(use_gy)) + reread + n_out * f * e * 4
                          + (e * 4 + 4 + f * 4 if use_gy else 0))
            _call("lookup_bwd_rows", lib.rsb_lookup_bwd_rows,
                  kind, L.ptr(rows), b, f, spec.dim, L.ptr(table), table.shape[0], L.ptr(table1), spec.divider,
                  L.ptr(aux_t), spec.modulus if spec.is_qr else spec.aux_mode, L.ptr(mask_d), L.ptr(emb), L.ptr(s),
                  L.ptr(g_y) if use_gy else None, L.ptr(g_emb), L.ptr(rg_main), L.ptr(rg_aux), None,
                  L.stream_ptr(dev), nbytes=nbytes)
            if kind == L.KIND_QR_ADD:
                rg_aux = rg_main

        # ---- stages 2+3: reduce by target row ---------------------------------
        g_table = g_table1 = g_aux = None
        n_rows = table.shape[0]
        mod = spec.module
        pre = ctx.presorted.get() if ctx.presorted is not None else None
        ctx.presorted = None
        deferred = getattr(mod, "_rsb_fused_opt", None) if (mod is not None and not spec.is_qr) else None
        # the lookups sorted by FULL row id (QR sorts by quotient): shared by the table, PEP-threshold and fc gradients
        full_pair = None if spec.is_qr else pre
        if full_pair is None and not spec.is_qr and need[4] and n > 0 and (deferred is not None or not spec.sparse_grad):
            full_pair = sort_rows(rows, n_rows)
        if want_fc:
            g_fc = fc_grad_beside(ctx.fc_param, rows, g_y, b, f, fc.shape, fc.shape[0], full_pair)
        if spec.is_qr:
            if need[4]:
                g_table = dense_row_grad(rows, rg_main, n_rows, key_div=spec.divider, sorted_pair=pre)
            if need[5] and g_table1_fused is not None:
                g_table1 = g_table1_fused
            elif need[5]:
                kmod = spec.modulus or spec.divider
                # <= 32 rows: register accumulators, no atomics; larger: float atomics in shared memory
                g_table1 = (small_table_grad(rows, rg_aux, table1.shape[0], key_mod=kmod)
                            if (not DETERMINISTIC or table1.shape[0] <= 32) else None)
                if g_table1 is None:
                    g_table1 = dense_row_grad(rows, rg_aux, table1.shape[0], key_mod=kmod)
        else:
            if need[4]:
                if deferred is not None:
                    deferred.stash(table, rows, rg_main, full_pair)     # consumed by FusedSparse*.step()
                elif spec.sparse_grad:
                    # same layout nn.Embedding(sparse=True) produces: uncoalesced COO, nnz = B*F
                    g_table = torch.sparse_coo_tensor(rows.view(1, n), rg_main, (n_rows, e))
                else:
                    pair = full_pair
                    g_table = dense_row_grad(rows, rg_main, n_rows, sorted_pair=pair)
                    if kind == L.KIND_PEP and need[6] and spec.aux_mode == L.PEP_FEATURE_DIM:
                        g_aux = dense_row_grad(rows, rg_aux, n_rows, sorted_pair=pair)
                    elif kind == L.KIND_PEP and need[6] and spec.aux_mode == L.PEP_FEATURE:
                        g_aux = dense_row_grad(rows, rg_aux.sum(dim=1, keepdim=True), n_rows, sorted_pair=pair)
            if kind == L.KIND_PEP and need[6] and g_aux is None:
                if spec.aux_mode == L.PEP_DIMENSION:
                    g_aux = rg_aux.sum(dim=0)
                elif spec.aux_mode == L.PEP_GLOBAL:
                    g_aux = rg_aux.sum().reshape(1)
                elif spec.aux_mode == L.PEP_FEATURE_DIM:
                    g_aux = dense_row_grad(rows, rg_aux, n_rows)
                else:
                    g_aux = dense_row_grad(rows, rg_aux.sum(dim=1, keepdim=True), n_rows)
            if kind == L.KIND_OPTEMBED and aux is not None and need[6]:
                g_aux = -rg_aux.sum(dim=0)
        return None, None, None, None, g_table, g_table1, g_aux, g_fc, g_bias, None, None


def fused_lookup(spec: LookupSpec, x: torch.Tensor, offsets: Optional[torch.Tensor], table: torch.Tensor,
                 table1: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
                 mask_d: Optional[torch.Tensor] = None, fc: Optional[torch.Tensor] = None,
                 bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Differentiable fused gather.  Returns (emb [B,VF,E], y_fm [B] or None)."""
    # The backward's row sort depends only on the looked-up rows: when a backward pass that reduces into the big table
    # will follow (grad mode on HERE, in the caller's context; a COO gradient needs no sort unless the fused row
    # optimizer consumes it), it is started on the side stream right after the forward gather (see EARLY_SORT).
    presort = (EARLY_SORT and torch.is_grad_enabled() and table.requires_grad and spec.kind != L.KIND_QR_CAT
               and not (spec.sparse_grad and getattr(spec.module, "_rsb_fused_opt", None) is None))
    slots = amax_slots_for(x.device)
    emb, y, _rows = _FusedLookup.apply(spec, x, offsets, mask_d, table, table1, aux, fc, bias, presort, slots)
    if slots is not None:
        emb._rsb_amax_slots = slots       # max(slots) == max |emb| (read by the dense tail's operand split)
    return emb, (y if fc is not None else None)


def csr_lookup(x: torch.Tensor, offsets: Optional[torch.Tensor], values: torch.Tensor, crow: torch.Tensor,
               col: torch.Tensor, n_rows: int, d: int, fc: Optional[torch.Tensor] = None,
               bias: Optional[torch.Tensor] = None, err_flag: Optional[torch.Tensor] = None):
    """Inference gather from a CSR-pruned table (pruned_embedding.py:89-173) fused with first order + FM.
    x [B,F] int32/int64 -> (emb [B,F,d], y_fm [B] or None).  No autograd: the reference class is inference-only."""
    lib = L.load()
    dev = L.require_cuda(x, values, crow, col, fc, bias, offsets)
    if x.dim() != 2:
        raise RuntimeError("rsb: x must be [B, F]")
    if x.dtype not in (torch.int32, torch.int64):
        raise RuntimeError(f"rsb: ids must be int32 or int64, got {x.dtype}")
    x = x.contiguous()
    b, f = x.shape
    emb = torch.empty(b, f, d, dtype=torch.float32, device=dev)
    y = torch.empty(b, dtype=torch.float32, device=dev) if fc is not None else None
    nnz = values.numel()
    nbytes = b * f * (x.element_size() + 2 * crow.element_size() + d * 4) + min(nnz, b * f * d) * (
        4 + col.element_size())
    _call("csr_lookup_fwd", lib.rsb_csr_lookup_fwd, L.ptr(x), int(x.dtype == torch.int32), L.ptr(offsets), b, f, d,
          L.ptr(values), L.ptr(crow), crow.element_size(), L.ptr(col), col.element_size(), n_rows, L.ptr(fc),
          L.ptr(bias), L.ptr(emb), L.ptr(y), L.ptr(err_flag), L.stream_ptr(dev), nbytes=nbytes)
    return emb, y


def dhe_encode(ids: torch.Tensor, prefix: int, slopes: torch.Tensor, bias: torch.Tensor, primes: torch.Tensor,
               m: int, small_operands: bool) -> torch.Tensor:
    """Universal-hash codes of `ids` (any shape) -> [*ids.shape, k] fp32 (dh_embedding.py:194-236), no grad."""
    lib = L.load()
    dev = L.require_cuda(ids, slopes, bias, primes)
    if ids.dtype not in (torch.int32, torch.int64):
        raise RuntimeError(f"rsb: ids must be int32 or int64, got {ids.dtype}")
    flat = ids.reshape(-1).contiguous()
    k = slopes.numel()
    out = torch.empty(flat.numel(), k, dtype=torch.float32, device=dev)
    _call("dhe_encode", lib.rsb_dhe_encode, L.ptr(flat), int(flat.dtype == torch.int32), flat.numel(), int(prefix),
          L.ptr(slopes), L.ptr(bias), L.ptr(primes), k, int(m), int(bool(small_operands)), L.ptr(out),
          L.stream_ptr(dev), nbytes=flat.numel() * (k * 4 + flat.element_size()))
    return out.reshape(*ids.shape, k)


# ----------------------------------------------------------------------------
# full-table helpers
# ----------------------------------------------------------------------------
class _SoftThresholdTable(torch.autograd.Function):
    """sign(w) * relu(|w| - sigmoid(s)) over a whole (small) table, differentiable in w and s
    (element-wise thresholds); one pass forward, one pass backward."""

    @staticmethod
    def forward(ctx, w, s):
        out, _ = pep_threshold_table(w.detach(), s.detach(), L.PEP_FEATURE_DIM)
        ctx.save_for_backward(w, s)
        return out

    @staticmethod
    def backward(ctx, g):
        w, s = ctx.saved_tensors
        lib = L.load()
        g = g.contiguous()
        gw = torch.empty_like(w)
        gs = torch.empty_like(w)
        _call("pep_dense_bwd", lib.rsb_pep_dense_bwd, L.ptr(w), L.ptr(s), L.PEP_FEATURE_DIM, w.shape[0], w.shape[1],
              L.ptr(g), L.ptr(gw), L.ptr(gs), L.stream_ptr(w.device))
        return gw, gs


def soft_threshold_table(w: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    return _SoftThresholdTable.apply(w, s)


def pep_threshold_table(weight, s, threshold_type: int, want_out=True, want_count=False):
    lib = L.load()
    dev = L.require_cuda(weight, s)
    out = torch.empty_like(weight) if want_out else None
    cnt = torch.zeros(1, dtype=torch.int64, device=dev) if want_count else None
    L.check(lib.rsb_pep_threshold_table(L.ptr(weight), L.ptr(s), threshold_type, weight.shape[0], weight.shape[1],
                                        L.ptr(out), L.ptr(cnt), L.stream_ptr(dev)), "pep_threshold_table")
    return out, cnt


def optembed_eval_weight(weight, t_row, mask_d_row, norm: int, want_out=True, want_count=False):
    lib = L.load()
    dev = L.require_cuda(weight, t_row, mask_d_row)
    out = torch.empty_like(weight) if want_out else None
    cnt = torch.zeros(1, dtype=torch.int64, device=dev) if want_count else None
    L.check(lib.rsb_optembed_eval_weight(L.ptr(weight), L.ptr(t_row), L.ptr(mask_d_row), norm, weight.shape[0],
                                         weight.shape[1], L.ptr(out), L.ptr(cnt), L.stream_ptr(dev)),
            "optembed_eval_weight")
    return out, cnt


def sigmoid(s: torch.Tensor) -> torch.Tensor:
    """sigmoid(s) in the kernels' own arithmetic (the threshold every PEP / CERP kernel compares against)."""
    lib = L.load()
    dev = L.require_cuda(s)
    s = s.contiguous()
    out = torch.empty_like(s)
    L.check(lib.rsb_sigmoid(L.ptr(s), s.numel(), L.ptr(out), L.stream_ptr(dev)), "sigmoid")
    return out


def mask_table(weight, mask):
    lib = L.load()
    dev = L.require_cuda(weight, mask)
    out = torch.empty_like(weight)
    m = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
    L.check(lib.rsb_mask_table(L.ptr(weight), L.ptr(m.contiguous()), weight.numel(), L.ptr(out), L.stream_ptr(dev)),
            "mask_table")
    return out
