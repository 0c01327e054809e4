"""CPU tests of the multi-GPU host logic: shard arithmetic (must agree with the kernels'
owner = row % G, local = row // G convention), full <-> shard state conversion, and the
world_size-2 flat gradient allreduce over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recsys_benchmark_b200 import sharded as S


@pytest.mark.parametrize("n,world", [(10, 2), (11, 2), (1086810, 8), (7, 4), (5, 8)])
def test_shard_round_trip(n, world):
    full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
    shards = [S.shard_of_full(full, r, world) for r in range(world)]
    assert all(s.shape == (S.shard_rows(n, world), 3) for s in shards)
    for row in range(n):
        g, l = S.owner_of(row, world), S.local_row(row, world)
        assert torch.equal(shards[g][l], full[row])
    assert torch.equal(S.full_from_shards(shards, n), full)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_round_trip_of_the_aux_arrays(world):
    """The per-row aux arrays of the sharded variants: a bool retrain mask [N, D] (moved as bytes) and PEP thresholds of
    the feature kind [N, 1]; padding rows of a shard are zero / False (a zero weight row survives no threshold)."""
    g = torch.Generator().manual_seed(world)
    n = 1373
    mask = torch.rand(n, 16, generator=g) > 0.4
    shards = [S.shard_of_full(mask, r, world) for r in range(world)]
    assert all(s.dtype == torch.bool and s.shape == (S.shard_rows(n, world), 16) for s in shards)
    for r, s in enumerate(shards):
        assert not bool(s[len(range(r, n, world)):].any())
    back = S.full_from_shards([s.view(torch.uint8) for s in shards], n).view(torch.bool)
    assert torch.equal(back, mask)
    col = torch.randn(n, 1, generator=g)
    assert torch.equal(S.full_from_shards([S.shard_of_full(col, r, world) for r in range(world)], n), col)


def test_hot_field_map_layout():
    """(lo, hi, delta) per field as the gather / scatter kernels read it: lo = the field offsets (ascending, so the
    scatter can binary-search a row's field), replicated fields packed in field order, hi == lo for sharded fields."""
    dims = [50, 7, 300, 11, 5, 1000, 3]
    m, h = S.hot_field_map(dims, 11)
    offs = np.concatenate([[0], np.cumsum(dims)[:-1]])
    assert m.dtype == torch.int64 and tuple(m.shape) == (7, 3) and h == 7 + 11 + 5 + 3
    np.testing.assert_array_equal(m[:, 0].numpy(), offs)
    hot = [d <= 11 for d in dims]
    base = 0
    for f, d in enumerate(dims):
        lo, hi, delta = m[f].tolist()
        assert hi - lo == (d if hot[f] else 0)
        if hot[f]:
            assert lo + delta == base
            base += d
    rows = S.hot_global_rows(m)
    assert rows.numel() == h
    # every replicated row maps to its slot through its own field's delta
    for slot, row in enumerate(rows.tolist()):
        f = int(np.searchsorted(offs, row, side="right") - 1)
        assert hot[f] and row + int(m[f, 2]) == slot
    m0, h0 = S.hot_field_map(dims, 0)
    assert h0 == 0 and torch.equal(m0[:, 0], m0[:, 1])
    mall, hall = S.hot_field_map(dims, 10 ** 9)
    assert hall == sum(dims) and torch.equal(S.hot_global_rows(mall), torch.arange(sum(dims)))
    import bench
    _, hc = S.hot_field_map(bench.CRITEO_DIMS, S.HOT_FIELD_ROWS)
    assert hc == sum(d for d in bench.CRITEO_DIMS if d <= 16384) and sum(d <= 16384 for d in bench.CRITEO_DIMS) == 31


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)
        grads = [torch.randn(5, 3), torch.randn(7), torch.randn(1)]
        ref = [g.clone() for g in grads]
        S.allreduce_mean_(grads)
        gathered = [None] * world
        dist.all_gather_object(gathered, [r.numpy() for r in ref])
        for i, g in enumerate(grads):
            mean = np.mean([gathered[r][i] for r in range(world)], axis=0)
            np.testing.assert_allclose(g.numpy(), mean, rtol=1e-6, atol=1e-7)
        # data-parallel invariant the sharded step relies on: mean of per-rank mean-loss grads
        # equals the grad of the global-mean loss
        w = torch.ones(3, requires_grad=True)
        xs = torch.arange(12, dtype=torch.float32).reshape(4, 3)
        local = xs[rank * 2:(rank + 1) * 2]
        (local @ w).mean().backward()
        g = [w.grad.clone()]
        S.allreduce_mean_(g)
        np.testing.assert_allclose(g[0].numpy(), xs.mean(0).numpy(), rtol=1e-6)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_world2_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}
