"""GPU tests of the row-sharded tables: world 1 (pointer-table path on a single GPU) and,
when the box has >= 2 GPUs, world 2 over NCCL + NVLink peer access against the single-GPU
model on the same global batch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

DIMS = [50, 7, 300, 11, 5, 1000, 3]


def _full_model(R, dev, seed=5):
    torch.manual_seed(seed)
    m = R.get_ctr_model(DIMS, dict(num_factor=16, hidden_sizes=[32, 16], p_dropout=0.0, use_batchnorm=False))
    return m.to(dev)


def _batch(b, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.stack([torch.randint(0, d, (b,), generator=g) for d in DIMS], 1)
    y = torch.randint(0, 2, (b,), generator=g).float()
    return x, y


def _sharded_from(R, full, dev, group=None, hot_field_rows=None):
    from recsys_benchmark_b200.sharded import ShardedDeepFM

    m = ShardedDeepFM(DIMS, 16, [32, 16], p_dropout=0.0, use_batchnorm=False, group=group,
                      embedding_config={"name": "vanilla", "hot_field_rows": hot_field_rows}).to(dev)
    hot = getattr(m.embedding._emb_module, "hot", None)
    assert (hot.shape[0] if hot is not None else 0) == sum(d for d in DIMS if d <= (hot_field_rows or 0))
    sg = m.embedding.shards
    st = {k: v for k, v in full.state_dict().items() if not k.startswith("embedding.")}
    m.load_state_dict(st, strict=False)            # fc.weight is replicated: same key / shape as the reference's
    m.embedding.load_full_weight(full.embedding.get_weight().detach())
    return m


def _train(model, opt, x, y, steps, sharded):
    crit = torch.nn.BCEWithLogitsLoss()
    outs = []
    for s in range(steps):
        logits = model(x)
        loss = crit(logits, y)
        opt.zero_grad()
        loss.backward()
        if sharded:
            model.sync_gradients()
        opt.step()
        if sharded:
            model.finish_step()
        outs.append(logits.detach().clone())
    return outs


@pytest.mark.parametrize("hot_field_rows", [None, 0, 11, 300, 10 ** 6])
def test_sharded_world1_matches_unsharded(hot_field_rows):
    """hot_field_rows: fields of at most that many ids are served from the replicated [H, D] table instead of the
    shard (None = the world-1 default: nothing replicated; 10**6: every field)."""
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200 as R

    dev = torch.device("cuda:0")
    full = _full_model(R, dev)
    sh = _sharded_from(R, full, dev, hot_field_rows=hot_field_rows)
    x, y = _batch(64, 1)
    x, y = x.to(dev), y.to(dev)
    o1 = torch.optim.Adam(full.parameters(), lr=1e-2, weight_decay=1e-4)
    o2 = torch.optim.Adam(sh.parameters(), lr=1e-2, weight_decay=1e-4)
    a = _train(full, o1, x, y, 3, False)
    b = _train(sh, o2, x, y, 3, True)
    assert torch.equal(a[0], b[0])   # same rows, same kernel arithmetic
    for s in range(3):
        assert_close(b[s].cpu().numpy(), a[s].cpu().numpy(), what=f"logits step {s}", atol_scale=5e-5)
    assert_close(sh.embedding.gather_full_weight().cpu().numpy(),
                 full.embedding.get_weight().detach().cpu().numpy(), what="table after 3 steps", atol_scale=5e-5)
    assert tuple(sh.fc.weight.shape) == tuple(full.fc.weight.shape)
    assert_close(sh.fc.weight.detach().cpu().numpy(), full.fc.weight.detach().cpu().numpy(),
                 what="fc after 3 steps", atol_scale=5e-5)


def test_shard_group_close_returns_the_ipc_buffers():
    """ShardGroup.close() frees the cudaMalloc'ed shard + gradient buffers (they do not come from torch's allocator)."""
    import __graft_entry__ as G

    G.build()
    from recsys_benchmark_b200.sharded import ShardGroup

    dev = torch.device("cuda:0")
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(dev)
    sg = ShardGroup(1 << 20, 16, dev, aux_cols=16, aux_grad=True)          # 4 buffers x 64 MiB
    sg.buf["table"].tensor.fill_(1.0)
    sg.zero_grads()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(dev)
    assert free0 - free1 >= 4 * (1 << 20) * 16 * 4 - (8 << 20)
    sg.close()
    free2, _ = torch.cuda.mem_get_info(dev)
    assert free2 - free1 >= 4 * (1 << 20) * 16 * 4 - (8 << 20)
    assert sg.buf["table"].tensor is None and sg.ptrs == {}
    sg.close()                                                              # idempotent


def _worker2(rank, world, port, ret, hot_field_rows):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import recsys_benchmark_b200 as R

        full = _full_model(R, dev)
        sh = _sharded_from(R, full, dev, hot_field_rows=hot_field_rows)
        x, y = _batch(128, 2)
        xl, yl = x[rank::world].to(dev), y[rank::world].to(dev)
        o1 = torch.optim.Adam(full.parameters(), lr=1e-2, weight_decay=1e-4)
        o2 = torch.optim.Adam(sh.parameters(), lr=1e-2, weight_decay=1e-4)
        a = _train(full, o1, x.to(dev), y.to(dev), 3, False)      # every rank: the single-GPU run on the global batch
        b = _train(sh, o2, xl, yl, 3, True)
        assert torch.equal(a[0][rank::world], b[0]), "forward over peer shards differs from the single-GPU gather"
        for s in range(3):
            assert_close(b[s].cpu().numpy(), a[s][rank::world].cpu().numpy(), what=f"logits step {s}", atol_scale=1e-4)
        assert_close(sh.embedding.gather_full_weight().cpu().numpy(),
                     full.embedding.get_weight().detach().cpu().numpy(), what="table", atol_scale=1e-4)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("hot_field_rows", [0, 50])
def test_sharded_world2_matches_single_gpu(hot_field_rows):
    """0: every field sharded; 50: the five small fields replicated, the 300- and 1000-id fields sharded."""
    import __graft_entry__ as G

    G.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker2, args=(2, 29700 + os.getpid() % 1000 + hot_field_rows, ret, hot_field_rows), nprocs=2,
             join=True)
    assert dict(ret) == {0: 1, 1: 1}
