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
    def __init__(self, loader: Iterable, device, depth: int = 2):
        self.loader = loader
        self.device = torch.device(device)
        self.depth = max(1, depth)
        self.stream = torch.cuda.Stream(self.device)

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch) -> Tuple[torch.Tensor, torch.Tensor, torch.cuda.Event]:
        x, y = batch
        if not x.is_pinned():
            x = x.pin_memory()
        if not y.is_pinned():
            y = y.pin_memory()
        with torch.cuda.stream(self.stream):
            xd = x.to(self.device, non_blocking=True)
            yd = y.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return xd, yd, ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        it = iter(self.loader)
        queue = []
        try:
            for _ in range(self.depth):
                queue.append(self._stage(next(it)))
        except StopIteration:
            pass
        while queue:
            xd, yd, ev = queue.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            xd.record_stream(cur)   # the copies were allocated on the side stream
            yd.record_stream(cur)
            try:
                queue.append(self._stage(next(it)))
            except StopIteration:
                pass
            yield xd, yd
