"""Input-side staging: the reference's loaders hand out int32 id batches
(`src/dataset/criteo/criteo_torchfm.py:72-93`, `avazu_fm.py:78-97`, `kdd_dataset.py:53-74`) and the
trainer copies them with a blocking `inputs.to(device)` right before the gather
(`src/trainer/deepfm.py:44-47`).  `DevicePrefetcher` wraps any iterable of (inputs, labels): batches are
staged in pinned host buffers and copied on a side stream one step ahead, so the H2D transfer of step
i+1 overlaps the compute of step i and `.to(device)` in the unchanged trainer loop is a no-op."""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import numpy as np
import torch

from . import _lib as L


class DevicePrefetcher:
    """Double-buffered: two persistent device buffers per tensor; the copy of batch i+1 is issued on a
    side stream after the compute stream has consumed batch i-1 (stream-ordered, no host sync), so it
    runs under step i.  No allocator traffic in steady state."""

    def __init__(self, loader: Iterable, device, depth: int = 2):
        self.loader = loader
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._bufs = [None, None]

    def __len__(self):
        return len(self.loader)

    def _buffer(self, slot: int, idx: int, like: torch.Tensor) -> torch.Tensor:
        cur = self._bufs[slot]
        if cur is None:
            cur = self._bufs[slot] = {}
        t = cur.get(idx)
        if t is None or t.shape != like.shape or t.dtype != like.dtype:
            t = cur[idx] = torch.empty(like.shape, dtype=like.dtype, device=self.device)
        return t

    def _stage(self, slot: int, batch):
        main = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(main)     # the buffer's previous consumer (two steps back) is done by then
        outs = []
        with torch.cuda.stream(self.stream):
            for idx, t in enumerate(batch):
                if not t.is_pinned():
                    t = t.pin_memory()
                dst = self._buffer(slot, idx, t)
                dst.copy_(t, non_blocking=True)
                outs.append(dst)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return outs, ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        it = iter(self.loader)
        try:
            nxt = self._stage(0, next(it))
        except StopIteration:
            return
        slot = 0
        while nxt is not None:
            outs, ev = nxt
            torch.cuda.current_stream(self.device).wait_event(ev)
            slot ^= 1
            try:
                nxt = self._stage(slot, next(it))
            except StopIteration:
                nxt = None
            yield tuple(outs)


class DeferredScalar:
    """Per-step device->host read of a scalar (the loss) that does not drain the launch queue.

    `loss.item()` right after `optimizer.step()` (src/trainer/deepfm.py:62) makes the host wait for the whole
    step before it can queue the next one, so every step starts with an empty GPU.  `push(loss)` copies the value
    into a pinned slot asynchronously and returns the value pushed ONE call earlier (None the first time): the
    D2H read still happens every step, the host just consumes it one step late; `flush()` returns the last one."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self._slots = torch.empty(depth, dtype=torch.float32).pin_memory()
        self._events = [torch.cuda.Event() for _ in range(depth)]
        self._n = 0

    def push(self, value: torch.Tensor):
        k = self._n % len(self._events)
        prev = None
        if self._n >= len(self._events) - 1 and self._n > 0:
            prev = self._read((self._n - 1) % len(self._events))
        self._slots[k:k + 1].copy_(value.detach().reshape(1), non_blocking=True)
        self._events[k].record(torch.cuda.current_stream(self.device))
        self._n += 1
        return prev

    def _read(self, k: int) -> float:
        self._events[k].synchronize()
        return float(self._slots[k])

    def flush(self):
        return self._read((self._n - 1) % len(self._events)) if self._n else None


def record_collate(samples):
    """`collate_fn` for the reference's lmdb datasets that keeps the cache's record layout: `__getitems__`
    (src/dataset/criteo/criteo_torchfm.py:79-93) reads one [B, F+1] block (label in column 0) and hands out row
    views `(arr[1:], arr[0])`; the default collate would re-stack them into two new arrays.  When the samples are
    views of one block the block itself is returned (zero copy), otherwise it is rebuilt.  -> int32 tensor [B, F+1]."""
    first = samples[0][0]
    base = getattr(first, "base", None)
    if (isinstance(base, np.ndarray) and base.ndim == 2 and base.shape == (len(samples), first.shape[0] + 1)
            and base.dtype in (np.int32, np.uint32) and base.flags.c_contiguous
            and all(getattr(x, "base", None) is base for x, _ in samples)
            and all(x.ctypes.data == base.ctypes.data + i * base.strides[0] + base.itemsize for i, (x, _) in enumerate(samples))):
        block = base
    else:
        block = np.empty((len(samples), len(first) + 1), dtype=np.int32)
        for i, (x, y) in enumerate(samples):
            block[i, 0] = int(y)
            block[i, 1:] = np.asarray(x)
    return torch.from_numpy(block.view(np.int32))


def unpack_records(records: torch.Tensor, ids_out: torch.Tensor = None, labels_out: torch.Tensor = None, stream=None):
    """records [B, F+1] int32 on the GPU (label, ids...) -> (ids [B,F] int32, labels [B] fp32): one `rsb_records_unpack`
    launch on `stream` (default: the current one)."""
    dev = L.require_cuda(records)
    if records.dtype != torch.int32 or records.dim() != 2 or records.shape[1] < 2 or not records.is_contiguous():
        raise ValueError("records must be a contiguous int32 [B, F+1] tensor")
    b, f = records.shape[0], records.shape[1] - 1
    ids = ids_out if ids_out is not None else torch.empty((b, f), dtype=torch.int32, device=dev)
    labels = labels_out if labels_out is not None else torch.empty((b,), dtype=torch.float32, device=dev)
    sp = stream.cuda_stream if stream is not None else L.stream_ptr(dev)
    L.check(L.load().rsb_records_unpack(L.ptr(records), b, f, L.ptr(ids), L.ptr(labels), sp), "rsb_records_unpack")
    return ids, labels


class RecordStager:
    """Record blocks [B, F+1] (what `record_collate` yields, or any int32 / uint32 array in the caches' layout) ->
    `(inputs int32 [B,F], labels fp32 [B])` already on the GPU: one pinned staging copy, ONE H2D transfer and the
    unpack kernel per batch, all on a side stream one step ahead of the compute stream.  The unchanged trainer's
    `inputs.to(device)`, `labels.to(device)` and `labels.float()` (src/trainer/deepfm.py:44-52) become no-ops.
    No bucket-by-owner pass exists for the row-sharded table: the gather reads peer rows in place (sharded.py)."""

    def __init__(self, loader: Iterable, device):
        self.loader = loader
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._slots = [None, None]

    def __len__(self):
        return len(self.loader)

    def _stage(self, slot: int, block):
        if isinstance(block, np.ndarray):
            block = torch.from_numpy(block.view(np.int32) if block.dtype == np.uint32 else block)
        if block.dtype != torch.int32 or block.dim() != 2:
            raise ValueError("RecordStager wants int32 / uint32 record blocks [B, F+1]")
        b, f1 = block.shape
        cur = self._slots[slot]
        if cur is None or cur["rec"].shape[0] < b or cur["rec"].shape[1] != f1:
            cur = self._slots[slot] = dict(pinned=None, rec=torch.empty((b, f1), dtype=torch.int32, device=self.device),
                                           ids=torch.empty((b, f1 - 1), dtype=torch.int32, device=self.device),
                                           labels=torch.empty((b,), dtype=torch.float32, device=self.device),
                                           copied=torch.cuda.Event(), src=None)
        cur["copied"].synchronize()            # the previous H2D copy out of this slot's host buffer has finished
        if block.is_pinned() and block.is_contiguous():
            src = cur["src"] = block           # e.g. DataLoader(pin_memory=True): copied from where it is, kept alive
        else:
            if cur["pinned"] is None or cur["pinned"].shape[0] < b or cur["pinned"].shape[1] != f1:
                cur["pinned"] = torch.empty((max(b, cur["rec"].shape[0]), f1), dtype=torch.int32).pin_memory()
            src = cur["pinned"][:b]
            src.copy_(block)
        main = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(main)          # the slot's previous consumer (two steps back) is done by then
        with torch.cuda.stream(self.stream):
            cur["rec"][:b].copy_(src, non_blocking=True)
            cur["copied"].record(self.stream)
            out = unpack_records(cur["rec"][:b], cur["ids"][:b], cur["labels"][:b], stream=self.stream)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return out, ready

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        it = iter(self.loader)
        try:
            nxt = self._stage(0, next(it))
        except StopIteration:
            return
        slot = 0
        while nxt is not None:
            out, ready = nxt
            torch.cuda.current_stream(self.device).wait_event(ready)
            slot ^= 1
            try:
                nxt = self._stage(slot, next(it))
            except StopIteration:
                nxt = None
            yield out
