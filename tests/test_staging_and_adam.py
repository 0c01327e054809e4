"""Record-format input staging and the one-launch dense Adam: host logic on the CPU, kernels on the GPU
against the numpy oracle, torch.optim.Adam and the plain reference loop."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import assert_close

DEV = "cuda:0"


def _records(b, field_dims, seed=0):
    rng = np.random.default_rng(seed)
    rec = np.empty((b, len(field_dims) + 1), dtype=np.uint32)
    rec[:, 0] = rng.integers(0, 2, b)
    for f, d in enumerate(field_dims):
        rec[:, 1 + f] = rng.integers(0, d, b)
    return rec


# ---------------------------------------------------------------- host logic (no GPU)
def test_record_collate_returns_the_block_itself_for_row_views():
    """`__getitems__` of the reference's lmdb datasets (src/dataset/criteo/criteo_torchfm.py:79-93) hands out row views
    of ONE int32 block: the collate gives that block back without copying."""
    from recsys_benchmark_b200.data import record_collate

    block = _records(17, [5, 9, 3]).astype(np.int32)
    samples = [(arr[1:], arr[0]) for arr in block]          # exactly what the dataset returns
    out = record_collate(samples)
    assert out.dtype == torch.int32 and tuple(out.shape) == (17, 4)
    assert out.data_ptr() == block.ctypes.data
    # shuffled / copied samples are rebuilt, same content
    perm = np.random.default_rng(1).permutation(17)
    out2 = record_collate([samples[i] for i in perm])
    assert out2.data_ptr() != block.ctypes.data
    np.testing.assert_array_equal(out2.numpy(), block[perm])
    # `__getitem__` style samples (int64 copies, criteo_torchfm.py:72-77)
    out3 = record_collate([(arr[1:].astype(np.int64), np.int64(arr[0])) for arr in block])
    np.testing.assert_array_equal(out3.numpy(), block)


def test_get_optimizers_selects_the_one_launch_adam_only_when_asked():
    import recsys_benchmark_b200 as R
    from recsys_benchmark_b200.optim import FusedDenseAdam

    model = R.get_ctr_model([7, 3, 11], dict(num_factor=8, hidden_sizes=[16], p_dropout=0.0))
    cfg = dict(learning_rate=1e-3, weight_decay=1e-6)
    assert type(R.get_optimizers(model, cfg)[0]) is torch.optim.Adam
    opt = R.get_optimizers(model, dict(cfg, fused_adam="rsb"))[0]
    assert isinstance(opt, FusedDenseAdam)
    assert opt.defaults["lr"] == 1e-3 and opt.defaults["weight_decay"] == 1e-6 and opt.defaults["betas"] == (0.9, 0.999)
    assert len(opt.param_groups[0]["params"]) == len(list(model.parameters()))
    sparse = R.get_optimizers(model, dict(cfg, sparse=True, fused_adam="rsb"))
    assert isinstance(sparse[0], torch.optim.SparseAdam) and isinstance(sparse[1], FusedDenseAdam)
    # CPU parameters are refused loudly (no fallback)
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError):
        opt.step()


def test_fused_dense_adam_launch_plan(monkeypatch):
    """Host side of the one-launch Adam without a GPU (the C call is recorded instead of made): one launch per distinct
    step count, the (tensor, chunk) map covers every element once, descriptors follow the gradient buffers, the plan is
    rebuilt when the set of parameters with gradients changes."""
    import ctypes as C

    from recsys_benchmark_b200 import _lib as L
    from recsys_benchmark_b200 import functional as RF
    from recsys_benchmark_b200.optim import FusedDenseAdam

    calls = []

    class FakeLib:
        @staticmethod
        def rsb_adam_dense(desc, n, bm_ptr, n_blocks, lr, b1, b2, eps, wd, step, stream):
            calls.append(dict(desc=[(d.param, d.grad, d.exp_avg, d.exp_avg_sq, d.numel) for d in desc], n=n,
                              n_blocks=n_blocks, lr=lr, wd=wd, step=step))
            return 0

    monkeypatch.setattr(L, "load", lambda: FakeLib)
    monkeypatch.setattr(L, "require_cuda", lambda *t: None)
    monkeypatch.setattr(L, "stream_ptr", lambda device=None: 0)
    monkeypatch.setattr(RF, "_TIMER", None)
    sizes = [1, 4096, 4097, 10000]
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
    opt = FusedDenseAdam(ps, lr=3e-4, weight_decay=1e-5)
    for p in ps:
        p.grad = torch.ones_like(p)
    opt.step()
    assert len(calls) == 1 and calls[0]["n"] == 4 and calls[0]["step"] == 1 and calls[0]["lr"] == 3e-4 and calls[0]["wd"] == 1e-5
    assert calls[0]["n_blocks"] == sum((n + L.ADAM_CHUNK - 1) // L.ADAM_CHUNK for n in sizes) == 1 + 1 + 2 + 3
    bm = next(iter(opt._maps.values()))
    assert bm.dtype == torch.int32 and bm.tolist() == [[0, 0], [1, 0], [2, 0], [2, 1], [3, 0], [3, 1], [3, 2]]
    for (pp, gp, mp, vp, n), p in zip(calls[0]["desc"], ps):
        assert (pp, gp, n) == (p.data_ptr(), p.grad.data_ptr(), p.numel())
        assert mp == opt.state[p]["exp_avg"].data_ptr() and vp == opt.state[p]["exp_avg_sq"].data_ptr()
    # new gradient buffers: same plan object, descriptors rewritten
    plan = opt._plans[0]
    for p in ps:
        p.grad = torch.ones_like(p)
    opt.step()
    assert opt._plans[0] is plan and calls[1]["step"] == 2
    assert [d[1] for d in calls[1]["desc"]] == [p.grad.data_ptr() for p in ps]
    # a parameter without a gradient is skipped; when it comes back it has its own step count -> its own launch
    ps[1].grad = None
    opt.step()
    assert calls[2]["n"] == 3 and calls[2]["step"] == 3 and opt.state[ps[1]]["step"] == 2
    ps[1].grad = torch.ones_like(ps[1])
    opt.step()
    assert sorted((c["step"], c["n"]) for c in calls[3:]) == [(3, 1), (4, 3)]
    assert [opt.state[p]["step"] for p in ps] == [4, 3, 4, 4]
    # gradients the kernel cannot take are refused
    ps[3].grad = torch.ones(20000)[::2]              # not contiguous
    with pytest.raises(RuntimeError, match="contiguous"):
        opt.step()


# ---------------------------------------------------------------- kernels
@pytest.mark.gpu
@pytest.mark.parametrize("b,dims", [(1, [4]), (257, [7, 3, 11, 5, 2, 9]), (4099, [1000] * 39), (65536, [50] * 11)])
def test_records_unpack_is_bit_exact(b, dims):
    import __graft_entry__ as G

    G.build()
    from recsys_benchmark_b200.data import unpack_records

    rec = _records(b, dims, seed=b)
    if b == 257:
        rec[3, 2] = 0x7fffffff                     # the largest id an int32 batch can carry
    ids, labels = unpack_records(torch.from_numpy(rec.view(np.int32)).to(DEV))
    np.testing.assert_array_equal(ids.cpu().numpy(), rec[:, 1:].astype(np.int32))
    np.testing.assert_array_equal(labels.cpu().numpy(), rec[:, 0].astype(np.float32))
    assert ids.dtype == torch.int32 and labels.dtype == torch.float32


@pytest.mark.gpu
def test_records_unpack_refuses_bad_blocks():
    from recsys_benchmark_b200.data import unpack_records

    with pytest.raises(ValueError):
        unpack_records(torch.zeros((4, 5), dtype=torch.int64, device=DEV))
    with pytest.raises(ValueError):
        unpack_records(torch.zeros((4, 6), dtype=torch.int32, device=DEV)[:, :5])
    with pytest.raises(RuntimeError):
        unpack_records(torch.zeros((4, 5), dtype=torch.int32))
    ids, labels = unpack_records(torch.zeros((0, 5), dtype=torch.int32, device=DEV))
    assert tuple(ids.shape) == (0, 4) and tuple(labels.shape) == (0,)


@pytest.mark.gpu
def test_record_stager_feeds_the_unchanged_loop_with_the_same_batches():
    """The staged path (records -> one H2D copy -> unpack kernel) and the reference's path (default collate ->
    inputs.to(device), labels.float()) give the trainer's step bit-identical inputs, hence bit-identical logits."""
    import recsys_benchmark_b200 as R
    from recsys_benchmark_b200.data import RecordStager

    dims = [50, 7, 300, 11, 5, 1000]
    blocks = [_records(n, dims, seed=s).view(np.int32) for s, n in enumerate([96, 96, 96, 40])]   # ragged tail
    torch.manual_seed(0)
    model = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[32], p_dropout=0.0)).to(DEV).eval()
    seen = 0
    with torch.no_grad():
        for (inputs, labels), blk in zip(RecordStager(blocks, DEV), blocks):
            inputs, labels = inputs.to(DEV), labels.to(DEV)          # src/trainer/deepfm.py:46-47: no-ops here
            assert labels.float() is labels
            ref_in = torch.from_numpy(blk[:, 1:].copy()).to(DEV)
            assert torch.equal(inputs, ref_in)
            assert torch.equal(labels, torch.from_numpy(blk[:, 0].copy()).to(DEV).float())
            assert torch.equal(model(inputs), model(ref_in))
            seen += 1
    assert seen == len(blocks)


@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 1e-6, 1e-2])
def test_fused_dense_adam_matches_torch_adam_and_the_oracle(wd):
    """Five steps over tensors of awkward sizes (1 element, a chunk boundary, a non-multiple of 4, a misaligned view)
    against torch.optim.Adam (what get_optimizers builds, src/models/deepfm.py:160-165) and the numpy oracle."""
    import __graft_entry__ as G

    G.build()
    from recsys_benchmark_b200.optim import FusedDenseAdam

    g = torch.Generator().manual_seed(3)
    base = torch.randn(4096 * 3 + 8, generator=g)
    shapes = [(1,), (4096,), (4097,), (1000, 16), (33, 7), (3,)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ours.append(torch.nn.Parameter(base.to(DEV)[1:4096 * 2 + 4]))              # 4-byte aligned only
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    host = [p.detach().cpu().numpy().copy() for p in ours]
    hm = [np.zeros_like(h) for h in host]
    hv = [np.zeros_like(h) for h in host]
    o1 = FusedDenseAdam(ours, lr=1e-3, weight_decay=wd)
    o2 = torch.optim.Adam(theirs, lr=1e-3, weight_decay=wd, foreach=False, fused=False)
    for step in range(1, 6):
        for i, (a, b) in enumerate(zip(ours, theirs)):
            gr = torch.randn(a.shape, generator=g) * (10.0 ** (i % 3 - 1))
            a.grad = gr.to(DEV)
            b.grad = gr.to(DEV)
            O.dense_adam(host[i], gr.numpy(), hm[i], hv[i], step, lr=1e-3, weight_decay=wd)
        o1.step()
        o2.step()
    for i, (a, b) in enumerate(zip(ours, theirs)):
        assert_close(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=1e-6, atol_scale=1e-7, what=f"param {i} vs torch")
        assert_close(a.detach().cpu().numpy(), host[i], rtol=2e-6, atol_scale=2e-7, what=f"param {i} vs oracle")
        assert_close(o1.state[a]["exp_avg"].cpu().numpy(), o2.state[b]["exp_avg"].cpu().numpy(), rtol=1e-6, atol_scale=1e-7,
                     what=f"exp_avg {i}")
        assert_close(o1.state[a]["exp_avg_sq"].cpu().numpy(), o2.state[b]["exp_avg_sq"].cpu().numpy(), rtol=1e-6,
                     atol_scale=1e-7, what=f"exp_avg_sq {i}")


@pytest.mark.gpu
def test_fused_dense_adam_skips_gradless_parameters_and_counts_steps_per_parameter():
    from recsys_benchmark_b200.optim import FusedDenseAdam

    g = torch.Generator().manual_seed(5)
    a, b = (torch.nn.Parameter(torch.randn(100, generator=g).to(DEV)) for _ in range(2))
    ta, tb = (torch.nn.Parameter(p.detach().clone()) for p in (a, b))
    o1 = FusedDenseAdam([a, b], lr=1e-2)
    o2 = torch.optim.Adam([ta, tb], lr=1e-2)
    for step in range(4):
        ga = torch.randn(100, generator=g).to(DEV)
        a.grad, ta.grad = ga, ga.clone()
        if step % 2:                                   # b only gets a gradient every other step
            gb = torch.randn(100, generator=g).to(DEV)
            b.grad, tb.grad = gb, gb.clone()
        else:
            b.grad = tb.grad = None
        o1.step()
        o2.step()
    assert o1.state[a]["step"] == 4 and o1.state[b]["step"] == 2
    assert_close(a.detach().cpu().numpy(), ta.detach().cpu().numpy(), rtol=1e-6, atol_scale=1e-7, what="a")
    assert_close(b.detach().cpu().numpy(), tb.detach().cpu().numpy(), rtol=1e-6, atol_scale=1e-7, what="b")


@pytest.mark.gpu
def test_deepfm_trains_the_same_with_the_one_launch_adam():
    """Whole model, 5 steps of the reference loop driven by get_optimizers(fused_adam="rsb"); a shadow copy of every
    parameter is stepped by the default torch Adam ON THE SAME GRADIENTS (feeding the two optimizers from two separately
    trained models would compare two chaotic trajectories: Adam turns a rounding-level gradient difference into +-lr)."""
    import recsys_benchmark_b200 as R

    dims = [50, 7, 300, 11, 5, 1000]
    cfg = dict(num_factor=16, hidden_sizes=[64, 64], p_dropout=0.0, use_batchnorm=True)
    torch.manual_seed(0)
    model = R.get_ctr_model(dims, dict(cfg)).to(DEV)
    opts = R.get_optimizers(model, dict(learning_rate=1e-3, weight_decay=1e-6, fused_adam="rsb"))
    named = [(n, p) for n, p in model.named_parameters()]
    shadow = [torch.nn.Parameter(p.detach().clone()) for _, p in named]
    ref = torch.optim.Adam(shadow, lr=1e-3, weight_decay=1e-6)
    crit = torch.nn.BCEWithLogitsLoss()
    for s in range(5):
        rec = _records(256, dims, seed=10 + s)
        x = torch.from_numpy(rec[:, 1:].astype(np.int32)).to(DEV)
        y = torch.from_numpy(rec[:, 0].astype(np.float32)).to(DEV)
        loss = crit(model(x), y)
        for o in opts:
            o.zero_grad()
        loss.backward()
        for (_, p), q in zip(named, shadow):
            q.grad = None if p.grad is None else p.grad.detach().clone()
        for o in opts:
            o.step()
        ref.step()
        for (n, p), q in zip(named, shadow):
            a, b = p.detach().cpu().numpy(), q.detach().cpu().numpy()
            # one step from the same state and gradient: rounding only (a few ulp of the update, which is <= lr)
            assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max() + 1e-3 * 1e-3, f"step {s} {n}: {np.abs(a - b).max():.3e}"
            q.data.copy_(p.detach())        # keep the two in the same state; the moments are compared at the end
    for (n, p), q in zip(named, shadow):
        if p.grad is None:
            continue
        st1, st2 = opts[0].state[p], ref.state[q]
        assert_close(st1["exp_avg"].cpu().numpy(), st2["exp_avg"].cpu().numpy(), rtol=1e-5, atol_scale=1e-6, what=f"exp_avg {n}")
        assert_close(st1["exp_avg_sq"].cpu().numpy(), st2["exp_avg_sq"].cpu().numpy(), rtol=1e-5, atol_scale=1e-6,
                     what=f"exp_avg_sq {n}")
