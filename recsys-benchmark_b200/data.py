"""Input-side staging: the reference's loaders hand out int32 id batches
(`src/dataset/criteo/criteo_torchfm.py:72-93`, `avazu_fm.py:78-97`, `kdd_dataset.py:53-74`) and the
trainer copies them with a blocking `inputs.to(device)` right before the gather
(`src/trainer/deepfm.py:44-47`).  `DevicePrefetcher` wraps any iterable of (inputs, labels): batches are
staged in pinned host buffers and copied on a side stream one step ahead, so the H2D transfer of step
i+1 overlaps the compute of step i and `.to(device)` in the unchanged trainer loop is a no-op."""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


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
