"""Row-sharded embedding tables over the GPUs of one box.

The reference is single-device; this is the scale-out of its ONE concatenated table
(`src/models/embeddings/base.py:53-57`, addressed by global row id, `src/models/deepfm.py:88`):
rank g owns rows r with r % G == g at local row r // G (block-cyclic: balances the very
uneven field sizes).  The batch is data parallel, dense parameters are replicated.

No index / row all-to-all is issued: shards live in CUDA-IPC-exported buffers and
  * the forward gather kernel reads peer rows directly over NVLink (rsb_lookup_fwd_sharded),
  * the backward sorts + pre-reduces this rank's lookups and adds each unique row's sum into
    the OWNER's dense shard gradient with 128-bit atomics over NVLink
    (rsb_segment_scatter_shards),
so the exchange is fused into the gather / scatter kernels.  Two stream-ordered NCCL
collectives per step provide the cross-rank ordering: the dense-gradient allreduce (all
pushes are complete once it returns) and a 1-element allreduce after the optimizer step
(shards updated and gradient buffers re-zeroed before anyone gathers again).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _lib as L
from . import functional as RF
from .deepfm import DeepFM
from .embeddings import IEmbedding
from .linalg import run_sequential


# ------------------------------------------------------------------ host-side shard math ---
def shard_rows(num_rows: int, world: int) -> int:
    """Rows per shard (uniform; the last rows of some shards may be padding)."""
    return (num_rows + world - 1) // world


def owner_of(row, world: int):
    return row % world


def local_row(row, world: int):
    return row // world


def shard_of_full(full: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows of `full` owned by `rank`, padded with zeros to shard_rows()."""
    n = shard_rows(full.shape[0], world)
    out = torch.zeros((n,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    mine = full[rank::world]
    out[: mine.shape[0]] = mine
    return out


def full_from_shards(shards: List[torch.Tensor], num_rows: int) -> torch.Tensor:
    world = len(shards)
    out = torch.empty((num_rows,) + tuple(shards[0].shape[1:]), dtype=shards[0].dtype, device=shards[0].device)
    for g, s in enumerate(shards):
        cnt = len(range(g, num_rows, world))
        out[g::world] = s[:cnt]
    return out


# Fields with at most this many ids are replicated on every rank instead of sharded (world > 1).  Criteo shape: 31 of
# the 39 fields (42 644 of 1 086 810 rows, 2.7 MB) - they serve 31 of a sample's 39 lookups, so the rows crossing NVLink
# drop from 39 * (G-1)/G to 8 * (G-1)/G per sample; their [H, D] gradient rides the replicated-gradient allreduce.
HOT_FIELD_ROWS = int(os.environ.get("RSB_HOT_FIELD_ROWS", "16384"))


def hot_field_map(field_dims, hot_field_rows: int):
    """-> (int64 [F, 3] rows (lo, hi, delta), H).  lo = the field's offset in the concatenated table; a replicated
    ("hot") field has hi = lo + dim and lives at rows [lo + delta, hi + delta) of the [H, D] replica; hi == lo marks
    a sharded field.  The layout rsb_lookup_fwd_sharded / rsb_segment_scatter_shards take."""
    rows, off, base = [], 0, 0
    for d in field_dims:
        d = int(d)
        hot = d <= hot_field_rows
        rows.append((off, off + d if hot else off, base - off if hot else 0))
        base += d if hot else 0
        off += d
    return torch.tensor(rows, dtype=torch.int64).reshape(-1, 3), base


def hot_global_rows(hot_map: torch.Tensor) -> torch.Tensor:
    """Global row ids of the replica's rows, in replica order."""
    parts = [torch.arange(int(lo), int(hi), dtype=torch.int64) for lo, hi, _ in hot_map.tolist() if hi > lo]
    return torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64)


def allreduce_mean_(grads: List[torch.Tensor], group=None) -> None:
    """Average a list of gradient tensors over the ranks with ONE flat allreduce (in place)."""
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    views, off = [], 0
    for g in grads:
        views.append(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    torch._foreach_copy_(grads, views)


# ------------------------------------------------------------------ IPC buffers ---
class SharedBuffer:
    """cudaMalloc'ed, zero-filled, CUDA-IPC-exportable fp32 buffer viewed as a torch tensor."""

    def __init__(self, shape, device: torch.device):
        lib = L.load()
        self.shape = tuple(int(s) for s in shape)
        self.device = device
        nbytes = 4 * max(1, math.prod(self.shape))
        p = C.c_void_p()
        with torch.cuda.device(device):
            L.check(lib.rsb_shared_alloc(nbytes, C.byref(p)), "shared_alloc")
        self.ptr = int(p.value)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False),
                                         "version": 3, "strides": None}
        self.tensor = torch.as_tensor(self, device=device)

    def handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            L.check(L.load().rsb_ipc_get_handle(self.ptr, buf), "ipc_get_handle")
        return buf.raw


def _open(handle: bytes, device) -> int:
    p = C.c_void_p()
    with torch.cuda.device(device):
        L.check(L.load().rsb_ipc_open_handle(handle, C.byref(p)), "ipc_open_handle")
    return int(p.value)


class ShardGroup:
    """The per-rank buffers (table shard and its gradient accumulator) and the device-resident pointer
    tables to every rank's copy."""

    NAMES = ("table", "table_grad")

    def __init__(self, num_rows: int, dim: int, device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.num_rows, self.dim, self.device = num_rows, dim, device
        self.n_local = shard_rows(num_rows, self.world)
        self.buf: Dict[str, SharedBuffer] = {
            "table": SharedBuffer((self.n_local, dim), device),
            "table_grad": SharedBuffer((self.n_local, dim), device)}
        handles = {k: self.buf[k].handle() for k in self.NAMES}
        if self.world > 1:
            gathered: List[Optional[dict]] = [None] * self.world
            dist.all_gather_object(gathered, handles, group=group)
        else:
            gathered = [handles]
        self.ptrs: Dict[str, torch.Tensor] = {}
        for k in self.NAMES:
            addr = [self.buf[k].ptr if g == self.rank else _open(gathered[g][k], device) for g in range(self.world)]
            self.ptrs[k] = torch.tensor(addr, dtype=torch.int64, device=device)
        self._tick = torch.zeros(1, device=device)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=device)   # set by the gather on an out-of-range id

    def barrier(self):
        """Stream-ordered cross-rank ordering point (a 1-element allreduce, no host sync)."""
        if self.world > 1:
            dist.all_reduce(self._tick, group=self.group)

    def zero_grads(self):
        self.buf["table_grad"].tensor.zero_()


# ------------------------------------------------------------------ differentiable op ---
class _ShardedLookup(torch.autograd.Function):
    """(x, offsets, fc, bias; shard group) -> emb [B,F,D], y_fm [B].  The table-shard gradients are not
    returned to autograd: they are accumulated (pre-scaled by 1/G) in the owners' buffers.  The first-order
    weights `fc` [N,1] are REPLICATED (4 B per row: peer reads / NVLink atomics of that size cost as many
    transactions as the 64-byte rows - at N=8 the sharded first-order gradient alone took 0.76 ms of a 5.0 ms
    step); their dense gradient goes back to autograd and is averaged with the other replicated gradients."""

    @staticmethod
    def forward(ctx, sg: ShardGroup, x, offsets, fc, bias, use_fm: bool, presort: bool = False, hot=None,
                hot_map=None, amax_slots=None):
        lib = L.load()
        dev = L.require_cuda(x, offsets, bias)
        x = x.contiguous()
        b, f = x.shape
        d = sg.dim
        emb = torch.empty(b, f, d, dtype=torch.float32, device=dev)
        y = torch.empty(b, dtype=torch.float32, device=dev) if use_fm else None
        s = torch.empty(b, d, dtype=torch.float32, device=dev) if use_fm else None
        rows = torch.empty(b, f, dtype=torch.int64, device=dev)
        if hot is not None and hot_map.shape[0] != f:
            raise ValueError(f"replicated small fields need one id column per field: x has {f} columns, the table "
                             f"{hot_map.shape[0]} fields")
        nbytes = b * (f * x.element_size() + 2 * f * d * 4 + f * 8 + (f * 4 + 4 + d * 4 if use_fm else 0))
        RF._call("lookup_fwd_sharded", lib.rsb_lookup_fwd_sharded, L.ptr(x), int(x.dtype == torch.int32),
                 L.ptr(offsets), b, f, d, L.ptr(sg.ptrs["table"]), L.ptr(fc) if use_fm else None,
                 sg.world, sg.num_rows, L.ptr(bias) if use_fm else None, L.ptr(hot) if hot is not None else None,
                 L.ptr(hot_map) if hot is not None else None, L.ptr(emb), L.ptr(y), L.ptr(s), L.ptr(rows),
                 L.ptr(sg.err_flag), L.ptr(amax_slots), L.stream_ptr(dev), nbytes=nbytes)
        ctx.sg, ctx.use_fm, ctx.shape = sg, use_fm, (b, f)
        ctx.hot_map = hot_map if hot is not None else None
        ctx.hot_shape = tuple(hot.shape) if hot is not None else None
        ctx.fc_shape = tuple(fc.shape) if fc is not None else None
        # the backward's row sort needs only `rows`: queued on the side stream now, consumed at the end of backward
        ctx.presorted = RF.early_sort(rows, sg.num_rows) if presort else None
        ctx.save_for_backward(rows, emb, s)
        ctx.mark_non_differentiable(rows)
        return emb, (y if use_fm else emb.new_empty(0)), rows

    @staticmethod
    def backward(ctx, g_emb, g_y, _g_rows):
        lib = L.load()
        sg: ShardGroup = ctx.sg
        rows, emb, s = ctx.saved_tensors
        b, f = ctx.shape
        d = sg.dim
        n = b * f
        dev = rows.device
        use_gy = ctx.use_fm and g_y is not None and g_y.numel() == b
        g_emb = g_emb.contiguous() if g_emb is not None else None
        if use_gy:
            g_y = g_y.contiguous()
            rg = torch.empty(n, d, dtype=torch.float32, device=dev)
            RF._call("lookup_bwd_rows", lib.rsb_lookup_bwd_rows, L.KIND_VANILLA, L.ptr(rows), b, f, d, L.ptr(emb),
                     sg.num_rows, None, 0, None, 0, None, L.ptr(emb), L.ptr(s), L.ptr(g_y), L.ptr(g_emb), L.ptr(rg),
                     None, None, L.stream_ptr(dev), nbytes=b * (f * 8 + 3 * f * d * 4 + d * 4 + 4))
        else:
            rg = g_emb.view(n, d)
        pre, ctx.presorted = ctx.presorted, None
        skeys, perm = pre.get() if pre is not None else RF.sort_rows(rows, sg.num_rows)
        scale = 1.0 / sg.world
        ws = RF._ws(lib.rsb_segment_workspace_bytes(n, d), dev)
        g_hot = None
        if ctx.hot_map is not None:
            # dense local gradient of the replicated rows (averaged over the ranks with the other replicated grads)
            g_hot = torch.zeros(ctx.hot_shape, dtype=torch.float32, device=dev)
        RF._call("segment_scatter_shards", lib.rsb_segment_scatter_shards, L.ptr(skeys), L.ptr(perm), n, L.ptr(rg), d,
                 L.ptr(sg.ptrs["table_grad"]), sg.world, scale,
                 L.ptr(ctx.hot_map) if g_hot is not None else None, f, L.ptr(g_hot) if g_hot is not None else None,
                 L.ptr(ws), ws.numel(), L.stream_ptr(dev), nbytes=n * (8 + 4 * d))
        g_bias = g_fc = None
        if use_gy:
            if ctx.needs_input_grad[3]:
                g_fc = RF.fc_grad(rows, g_y, b, f, ctx.fc_shape, sg.num_rows, (skeys, perm))
            g_bias = g_y.sum().reshape(1)
        return None, None, None, g_fc, g_bias, None, None, g_hot, None, None


# ------------------------------------------------------------------ modules ---
class ShardedVanillaEmbedding(IEmbedding):
    """VanillaEmbedding (base.py:23-75) whose table is row-sharded over the process group.
    `_emb_module.weight` is THIS rank's shard [ceil(N/G), D]; use `load_full_weight` /
    `gather_full_weight` to convert from / to the reference's full-table state dict."""

    def __init__(self, field_dims, hidden_size: int, mode=None, initializer="xavier", device=None, group=None,
                 hot_field_rows: Optional[int] = None, **kwargs):
        """hot_field_rows: fields of at most this many ids are replicated (`_emb_module.hot` [H, D], a plain
        replicated parameter) instead of sharded; default HOT_FIELD_ROWS when world > 1, nothing at world 1."""
        super().__init__()
        assert mode is None, "sharded tables serve the [B,F,D] lookup only"
        field_dims = [field_dims] if isinstance(field_dims, int) else [int(v) for v in field_dims]
        self._num_item = sum(field_dims)
        self._hidden_size = hidden_size
        device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.shards = ShardGroup(self._num_item, hidden_size, device, group)
        self._emb_module = nn.Module()
        self._emb_module.weight = nn.Parameter(self.shards.buf["table"].tensor)
        if hot_field_rows is None:
            hot_field_rows = HOT_FIELD_ROWS if self.shards.world > 1 else 0
        hot_map, n_hot = hot_field_map(field_dims, int(hot_field_rows))
        self.hot_map = hot_map.to(device) if n_hot > 0 else None
        self._hot_rows = hot_global_rows(hot_map).to(device) if n_hot > 0 else None
        if n_hot > 0:
            self._emb_module.hot = nn.Parameter(torch.zeros(n_hot, hidden_size, dtype=torch.float32, device=device))
        self._init_shard(initializer)

    INIT_CHUNK_ROWS = 1 << 20

    @torch.no_grad()
    def _init_shard(self, initializer: str):
        """i.i.d. init of the FULL table (base.py:62-67: xavier_uniform_ / normal_(std=0.1)), of which this rank keeps
        its rows: every rank draws the same global row blocks from the device generator (same seed on every rank, as
        the replicated parameters already require) and copies out rows r with r % G == rank.  The gathered table is
        the same for every world size, and no two global rows share values (drawing the local shard directly from
        the same seed would give rows kG..kG+G-1 identical vectors)."""
        sg = self.shards
        n, d, g = self._num_item, self._hidden_size, sg.world
        a = math.sqrt(6.0 / (n + d))            # xavier_uniform_ bound of the FULL [N, D] table
        w = self._emb_module.weight
        w.zero_()
        for lo in range(0, n, self.INIT_CHUNK_ROWS):
            hi = min(n, lo + self.INIT_CHUNK_ROWS)
            blk = torch.empty(hi - lo, d, dtype=torch.float32, device=sg.device)
            if initializer == "xavier":
                blk.uniform_(-a, a)
            else:
                blk.normal_(std=0.1)
            first = lo + ((sg.rank - lo) % g)   # first global row >= lo owned by this rank
            if first < hi:
                mine = blk[first - lo::g]
                w[first // g: first // g + mine.shape[0]].copy_(mine)
            if self._hot_rows is not None:
                sel = (self._hot_rows >= lo) & (self._hot_rows < hi)
                if bool(sel.any()):
                    self._emb_module.hot[sel] = blk[self._hot_rows[sel] - lo]

    def get_weight(self):
        return self.gather_full_weight()

    def load_full_weight(self, full: torch.Tensor):
        sg = self.shards
        with torch.no_grad():
            full = full.to(sg.device)
            self._emb_module.weight.copy_(shard_of_full(full, sg.rank, sg.world))
            if self._hot_rows is not None:
                self._emb_module.hot.copy_(full[self._hot_rows])

    def gather_full_weight(self) -> torch.Tensor:
        sg = self.shards
        local = self._emb_module.weight.detach()
        if sg.world == 1:
            full = local[: self._num_item].clone()
        else:
            parts = [torch.empty_like(local) for _ in range(sg.world)]
            dist.all_gather(parts, local.contiguous(), group=sg.group)
            full = full_from_shards(parts, self._num_item)
        if self._hot_rows is not None:
            full[self._hot_rows] = self._emb_module.hot.detach()   # the owners' copies of replicated rows are stale
        return full

    def lookup(self, x, offsets=None, fc=None, bias=None):
        if offsets is not None:
            offsets = offsets.reshape(-1).long()
        use_fm = fc is not None
        self._rsb_err_flag = self.shards.err_flag     # read by IEmbedding.train() / validate=True like the others
        presort = RF.EARLY_SORT and torch.is_grad_enabled()
        hot = getattr(self._emb_module, "hot", None)
        slots = RF.amax_slots_for(x.device)
        emb, y, _ = _ShardedLookup.apply(self.shards, x, offsets, fc, bias, use_fm, presort, hot, self.hot_map, slots)
        if slots is not None:
            emb._rsb_amax_slots = slots
        if self.validate:
            RF.check_index_errors(self)
        return emb, (y if use_fm else None)


class ShardedDeepFM(DeepFM):
    """DeepFM (src/models/deepfm.py:11-105) with the embedding table row-sharded; the first-order weights
    `fc` (4 B per row), the MLP, `_bias` and BatchNorm are replicated (BatchNorm statistics stay per-rank).
    Build it after `dist.init_process_group` with the device current and the same seed on every rank."""

    def __init__(self, field_dims, num_factor, hidden_sizes, p_dropout=0.1, use_batchnorm=False,
                 embedding_config=None, group=None):
        cfg = dict(embedding_config or {"name": "vanilla"})
        if cfg.get("name", "vanilla") != "vanilla":
            raise NotImplementedError("row sharding covers the full (vanilla) table; compressed tables are replicated")
        super().__init__(field_dims, num_factor, hidden_sizes, p_dropout, use_batchnorm, {"name": "vanilla"},
                         empty_embedding=True)
        cfg.pop("name", None)
        cfg.pop("sparse", None)
        self.embedding = ShardedVanillaEmbedding(field_dims, num_factor, group=group, **cfg)

    def shard_params(self):
        return [self.embedding._emb_module.weight]

    def replicated_params(self):
        ids = {id(p) for p in self.shard_params()}
        return [p for p in self.parameters() if id(p) not in ids and p.requires_grad]

    def forward(self, x):
        emb, y_fm = self.embedding.lookup(x, self.offsets, self.fc.weight, self._bias)
        b = emb.shape[0]
        scores = y_fm.unsqueeze(1) + run_sequential(self._deep_branch, emb.reshape(b, emb.shape[1] * emb.shape[2]),
                                                    overlap_first_dw=True, x_amax_slots=RF.amax_slots_of(emb))
        return scores.squeeze(-1)

    # -- step protocol ---------------------------------------------------------------
    def sync_gradients(self):
        """Call after backward(): averages the replicated gradients over the ranks (this
        allreduce is also the point after which every rank's pushes into our shard gradient
        are complete) and exposes the accumulated shard gradients as `.grad`."""
        sg = self.embedding.shards
        allreduce_mean_([p.grad for p in self.replicated_params() if p.grad is not None], sg.group)
        self.embedding._emb_module.weight.grad = sg.buf["table_grad"].tensor

    def finish_step(self):
        """Call after optimizer.step(): re-zero the shard gradient accumulators and order the
        next step's peer gathers / pushes after every rank's update."""
        sg = self.embedding.shards
        self.embedding._emb_module.weight.grad = None
        sg.zero_grads()
        sg.barrier()


class ShardedStepOptimizers:
    """Facade with the `zero_grad()` / `step()` interface the reference trainer drives
    (src/trainer/deepfm.py:52-60): step() = sync_gradients -> inner optimizers -> finish_step."""

    def __init__(self, model: ShardedDeepFM, optimizers: List[torch.optim.Optimizer]):
        self.model, self.optimizers = model, optimizers

    def zero_grad(self, set_to_none: bool = True):
        for p in self.model.replicated_params():
            p.grad = None

    def step(self):
        self.model.sync_gradients()
        for o in self.optimizers:
            o.step()
        self.model.finish_step()
