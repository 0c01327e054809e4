"""Deep hash embedding behind the reference's `DHEmbedding` API
(src/models/embeddings/dh_embedding.py:16-362): same constructor, buffers `_slopes` / `_bias` /
`_primes_choices`, `_seq` encoder (state-dict compatible), class-level `COUNTER` prefix, extra state.

B200-native differences, none of them visible in the numbers:
  * the [N, k] code table is never stored: `rsb_dhe_encode` regenerates the universal-hash code of every
    looked-up id inside the kernel, bit-exactly (the reference caches 4.4 GB at Criteo shape, k = 1024);
    `cached` / `cache_path` are accepted, `_cache` is materialised only if somebody reads it;
  * the encoder's Linear layers (the real GEMMs: [B*F, 1024] x [1024, h]) run on the fp32-accurate
    tensor-core kernel through `linalg.run_sequential`; BatchNorm1d / Mish are torch ops;
  * the prime table (src/assets/large_prime_74518.json = the first 74 518 primes above 10^6) is sieved at
    construction instead of shipped.
Only the universal hash (`use_universal_hash=True`, the reference default) is provided: the legacy variant
draws its coefficients from a per-item `torch.manual_seed(item)` CPU stream and cannot run on the device.
"""
from __future__ import annotations

import json
from typing import List, Optional, Union

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import functional as RF
from .embeddings import IEmbedding
from .linalg import run_sequential

LARGE_INT = int(1e9)
NEGATIVE_LARGE_INT = -LARGE_INT
_PRIME_CACHE = {}


def first_primes_above(lo: int, count: int) -> List[int]:
    key = (lo, count)
    if key not in _PRIME_CACHE:
        hi = int(lo * 2.2) + 1000
        while True:
            sieve = np.ones(hi + 1, dtype=bool)
            sieve[:2] = False
            for i in range(2, int(hi ** 0.5) + 1):
                if sieve[i]:
                    sieve[i * i::i] = False
            pr = np.nonzero(sieve)[0]
            pr = pr[pr > lo]
            if len(pr) >= count:
                _PRIME_CACHE[key] = pr[:count].tolist()
                break
            hi *= 2
    return _PRIME_CACHE[key]


class DHEmbedding(IEmbedding):
    COUNTER = 0

    def __init__(self, field_dims: Union[int, List[int]], out_size: int, mode: Optional[str] = None,
                 inp_size: int = 1024, hidden_sizes: Optional[List[int]] = None, use_bn: Union[bool, int] = 2,
                 cached: bool = True, prime_file: Optional[str] = None, cache_path: str = "", compute_v2=False,
                 use_universal_hash=True):
        super().__init__()
        if not use_universal_hash:
            raise NotImplementedError("DHEmbedding: only use_universal_hash=True runs on the device "
                                      "(the legacy hash seeds a CPU RNG stream per item)")
        if prime_file is None:
            primes = first_primes_above(10 ** 6, 74518)
        else:
            with open(prime_file) as fin:
                primes = json.load(fin)
        if isinstance(field_dims, int):
            field_dims = [field_dims]
        if isinstance(use_bn, bool):
            use_bn = int(use_bn)
        num_item = sum(field_dims)
        self.m = int(1e6)
        # prefix keeps user / item tables of one model apart (dh_embedding.py:74-77)
        self._prefix = DHEmbedding.COUNTER
        DHEmbedding.COUNTER += num_item

        self._primes = torch.tensor(primes)
        self._inp_size = inp_size
        self._num_item = num_item
        self._use_universal_hash = True
        rng = torch.Generator()
        rng.manual_seed(0)
        self.register_buffer("_slopes", self._random_nonzero_int(inp_size, rng))
        self.register_buffer("_bias", self._random_nonzero_int(inp_size, rng))
        p_idx = torch.randint(0, len(primes), (inp_size,), generator=rng)
        self.register_buffer("_primes_choices", self._primes[p_idx])

        layers: List[nn.Module] = []
        sizes = list(hidden_sizes) if hidden_sizes is not None else []
        sizes.append(out_size)
        for size in sizes:
            layers.append(nn.Linear(inp_size, size))
            if use_bn == 1:
                layers.append(nn.Mish())
                layers.append(nn.BatchNorm1d(size))
            elif use_bn == 2:
                layers.append(nn.BatchNorm1d(size))
                layers.append(nn.Mish())
            else:
                layers.append(nn.Mish())
            inp_size = size
        self._seq = nn.Sequential(*layers)
        self._use_cache = cached
        self._use_bn = use_bn
        self._out_size = out_size
        self._hidden_size = out_size
        self.compute_v2 = compute_v2
        self._mode = mode
        self._emb = None
        self._small = None      # lazily: do the operands fit the fast modulo?  (see rsb_dhe_encode)

    @staticmethod
    def _random_nonzero_int(num_element, rng=None):
        b = torch.randint(NEGATIVE_LARGE_INT, LARGE_INT, (num_element,), generator=rng)
        mask = b == 0
        while mask.sum() > 0:
            b[mask] = torch.randint(NEGATIVE_LARGE_INT, LARGE_INT, (int(mask.sum().item()),), generator=rng)
            mask = b == 0
        return b

    # ---- hash codes -------------------------------------------------------------------------------
    def _load_from_state_dict(self, *args, **kwargs):
        self._small = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _small_operands(self) -> bool:
        if self._small is None:
            lim = 2 ** 31
            self._small = bool(int(self._slopes.abs().max()) < lim and int(self._bias.abs().max()) < lim
                               and 2 <= int(self._primes_choices.min()) and int(self._primes_choices.max()) < lim)
        return self._small

    def encode(self, ids: torch.Tensor, in_table: bool = False) -> torch.Tensor:
        """Universal-hash codes [*ids.shape, k] of arbitrary ids (`_get_universal_hash_batch`,
        dh_embedding.py:215-236).  in_table=True promises 0 <= ids < num_item (fast modulo)."""
        small = in_table and self._small_operands() and self._num_item + self._prefix + 1 < 2 ** 30
        dev = ids.device   # the reference moves its coefficients to the ids' device too (dh_embedding.py:217-219)
        return RF.dhe_encode(ids, self._prefix, self._slopes.to(dev), self._bias.to(dev),
                             self._primes_choices.to(dev), self.m, small)

    _get_universal_hash_batch = encode

    @property
    def _cache(self) -> torch.Tensor:
        """The reference's cached code table [num_item, k] (dh_embedding.py:250-268), built on demand - on "cuda"
        whenever one is available, like the reference, wherever the module itself lives."""
        dev = self._slopes.device
        if dev.type != "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("rsb: DHEmbedding codes are generated on the GPU (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device())
        return self.encode(torch.arange(self._num_item, device=dev), in_table=True)

    # ---- forward ----------------------------------------------------------------------------------
    def get_weight(self):
        return self(torch.arange(self._num_item, device=self._seq[0].weight.device))

    def _forward_mlp(self, embs):
        is_flatten = embs.dim() == 3
        if is_flatten:
            batch, num_field, dimension = embs.shape
            embs = embs.reshape(batch * num_field, dimension)
        outs = run_sequential(self._seq, embs)
        if is_flatten:
            outs = outs.reshape(batch, num_field, -1)
        return outs

    def forward(self, inp: torch.Tensor):
        mode = self._mode
        if not self.training and self._emb is not None:
            return F.embedding(inp, self._emb)
        if self.compute_v2 and self._use_cache:
            uniques, inverse_idx = inp.unique(return_inverse=True)
            x = self._forward_mlp(self.encode(uniques, in_table=True))
            return x[inverse_idx]
        if self._use_cache:
            embs = self.encode(inp, in_table=True)
            if mode is not None:                       # F.embedding_bag(inp, cache, mode=mode)
                embs = embs.sum(1) if mode == "sum" else embs.mean(1) if mode == "mean" else embs.max(1).values
            return self._forward_mlp(embs)
        if self.training:
            raise NotImplementedError()                # reference: un-cached mode is inference-only
        uniques, inverse_idx = inp.unique(return_inverse=True)
        feats = run_sequential(self._seq, self.encode(uniques))
        return F.embedding(inverse_idx, feats)

    def lookup(self, x: torch.Tensor, offsets: Optional[torch.Tensor] = None, fc: Optional[torch.Tensor] = None,
               bias: Optional[torch.Tensor] = None):
        """(emb, y_fm) like the table plugins' fused entry; the embedding comes out of the encoder GEMMs, so
        first order + FM second order are element-wise passes over it (src/models/deepfm.py:88-98)."""
        rows = x.long()
        if offsets is not None:
            rows = rows + offsets.reshape(1, -1)
        emb = self(rows)
        if fc is None:
            return emb, None
        first = F.embedding_bag(rows, fc, mode="sum")
        if bias is not None:
            first = first + bias
        y = first + 0.5 * (emb.sum(dim=1).pow(2) - emb.pow(2).sum(dim=1)).sum(1, keepdim=True)
        return emb, y.squeeze(1)

    def set_extra_state(self, state):
        self._prefix = state["_prefix"]

    def get_extra_state(self):
        return {"_prefix": self._prefix}
