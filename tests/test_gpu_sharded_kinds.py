"""Row sharding of the lightweight variants (QR emb2, PEP weight + thresholds, retrain weight + mask).

Kernel level: three "virtual" shards on ONE GPU (the pointer table simply points at three tensors) must give the
single-table kernels' results bit for bit - forward, chain-rule backward and the pushes into the owners' gradient
shards.  Module level: ShardedDeepFM / ShardedDCNMix at world 1 against the single-device model, and at world 2 over
NCCL + NVLink peer access when the box has two GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
DIMS = [50, 7, 300, 11, 5, 1000, 3]
D = 16


@pytest.fixture(scope="module")
def env():
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200.functional as RF
    from recsys_benchmark_b200 import _lib as L
    from recsys_benchmark_b200 import sharded as S

    return RF, L, S


def _ids(b, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randint(0, d, (b,), generator=g) for d in DIMS], 1)


def _ptr_table(tensors):
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=DEV)


KIND_CASES = ["vanilla", "qr_mult", "qr_add", "pep_global", "pep_dimension", "pep_feature", "pep_feature_dim", "mask"]


@pytest.mark.parametrize("case", KIND_CASES)
@pytest.mark.parametrize("G", [1, 3])
def test_virtual_shards_match_the_single_table_kernels_bit_for_bit(env, case, G):
    RF, L, S = env
    lib = L.load()
    torch.manual_seed(7)
    n = sum(DIMS)
    b, f = 97, len(DIMS)
    x = _ids(b, 3).to(DEV)
    offsets = torch.tensor([0] + DIMS[:-1]).cumsum(0).to(DEV)
    kind = {"vanilla": L.KIND_VANILLA, "qr_mult": L.KIND_QR_MULT, "qr_add": L.KIND_QR_ADD, "mask": L.KIND_MASK}.get(case, L.KIND_PEP)
    divider = 7 if case.startswith("qr") else 0
    n_main = (n - 1) // divider + 1 if divider else n
    table = (torch.rand(n_main, D, device=DEV) - 0.5)
    table1 = torch.rand(divider, D, device=DEV) + 0.5 if divider else None
    aux, aux_mode, per_row = None, 0, False
    if kind == L.KIND_PEP:
        aux_mode = L.PEP_TYPES[case[4:]]
        shape = {"global": (1,), "dimension": (D,), "feature": (n, 1), "feature_dim": (n, D)}[case[4:]]
        aux = torch.randn(shape, device=DEV) - 1.5                      # sigmoid(s) ~ 0.2: about half the weights survive
        per_row = aux_mode in (L.PEP_FEATURE, L.PEP_FEATURE_DIM)
    elif kind == L.KIND_MASK:
        aux = (torch.rand(n, D, device=DEV) > 0.4).view(torch.uint8)
        per_row = True
    fc = torch.randn(n, 1, device=DEV)
    bias = torch.randn(1, device=DEV)

    def outputs():
        return (torch.empty(b, f, D, device=DEV), torch.empty(b, device=DEV), torch.empty(b, D, device=DEV),
                torch.empty(b, f, dtype=torch.int64, device=DEV))

    st = L.stream_ptr(torch.device(DEV))
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    emb0, y0, s0, rows0 = outputs()
    L.check(lib.rsb_lookup_fwd(kind, L.ptr(x), 0, L.ptr(offsets), b, f, D, L.ptr(table), n_main, n, L.ptr(table1), divider,
                               L.ptr(aux), aux_mode, None, L.ptr(fc), L.ptr(bias), L.ptr(emb0), L.ptr(y0), L.ptr(s0),
                               L.ptr(rows0), L.ptr(err), None, st))
    shards = [S.shard_of_full(table, g, G).contiguous() for g in range(G)]
    aux_shards = [S.shard_of_full(aux, g, G).contiguous() for g in range(G)] if per_row else None
    tp = _ptr_table(shards)
    ap = _ptr_table(aux_shards) if per_row else None
    emb1, y1, s1, rows1 = outputs()
    L.check(lib.rsb_lookup_fwd_sharded_kind(kind, L.ptr(x), 0, L.ptr(offsets), b, f, D, L.ptr(tp), G, n_main, n,
                                            L.ptr(table1), divider, None if per_row else L.ptr(aux), L.ptr(ap), aux_mode,
                                            None, L.ptr(fc), L.ptr(bias), L.ptr(emb1), L.ptr(y1), L.ptr(s1), L.ptr(rows1),
                                            L.ptr(err), None, st))
    assert int(err.item()) == 0
    assert torch.equal(rows0, rows1) and torch.equal(emb0, emb1) and torch.equal(y0, y1) and torch.equal(s0, s1)

    # chain-rule backward: per-lookup row gradients
    g_deep = torch.randn(b, f, D, device=DEV)
    g_y = torch.randn(b, device=DEV)
    two = kind in (L.KIND_QR_MULT, L.KIND_PEP)
    rg0, ra0 = torch.zeros(b * f, D, device=DEV), (torch.zeros(b * f, D, device=DEV) if two else None)
    rg1, ra1 = torch.zeros(b * f, D, device=DEV), (torch.zeros(b * f, D, device=DEV) if two else None)
    L.check(lib.rsb_lookup_bwd_rows(kind, L.ptr(rows0), b, f, D, L.ptr(table), n_main, L.ptr(table1), divider, L.ptr(aux),
                                    aux_mode, None, L.ptr(emb0), L.ptr(s0), L.ptr(g_y), L.ptr(g_deep), L.ptr(rg0),
                                    L.ptr(ra0), None, st))
    L.check(lib.rsb_lookup_bwd_rows_sharded(kind, L.ptr(rows0), b, f, D, L.ptr(tp), G, n_main, L.ptr(table1), divider,
                                            None if per_row else L.ptr(aux), L.ptr(ap), aux_mode, None, L.ptr(emb0),
                                            L.ptr(s0), L.ptr(g_y), L.ptr(g_deep), L.ptr(rg1), L.ptr(ra1), st))
    assert torch.equal(rg0, rg1)
    if two:
        assert torch.equal(ra0, ra1)

    # pushes into the owners' gradient shards == the dense scatter-add, re-assembled
    pair = RF.sort_rows(rows0, n_main, key_div=divider)
    dense = RF.dense_row_grad(rows0, rg0, n_main, key_div=divider, sorted_pair=pair)
    gshards = [torch.zeros_like(t) for t in shards]
    gp = _ptr_table(gshards)
    ws = RF._ws(lib.rsb_segment_workspace_bytes(b * f, D), torch.device(DEV))
    L.check(lib.rsb_segment_scatter_shards(L.ptr(pair[0]), L.ptr(pair[1]), b * f, L.ptr(rg0), D, L.ptr(gp), G, 1.0, None, f,
                                           None, L.ptr(ws), ws.numel(), st))
    assert torch.equal(S.full_from_shards(gshards, n_main), dense)       # one add per unique row into zeros: exact
    if case == "pep_feature":                                            # 1-wide rows (s [N,1]) pushed the same way
        col = ra0.sum(dim=1, keepdim=True).contiguous()
        dense1 = RF.dense_row_grad(rows0, col, n, sorted_pair=pair)
        g1 = [torch.zeros(t.shape[0], 1, device=DEV) for t in shards]
        ws = RF._ws(lib.rsb_segment_workspace_bytes(b * f, 1), torch.device(DEV))
        L.check(lib.rsb_segment_scatter_shards(L.ptr(pair[0]), L.ptr(pair[1]), b * f, L.ptr(col), 1, L.ptr(_ptr_table(g1)), G,
                                               1.0, None, f, None, L.ptr(ws), ws.numel(), st))
        assert torch.equal(S.full_from_shards(g1, n), dense1)


def test_sharded_kind_entry_points_refuse_inconsistent_arguments(env):
    RF, L, S = env
    lib = L.load()
    t = torch.zeros(8, D, device=DEV)
    tp = _ptr_table([t])
    x = torch.zeros(2, 1, dtype=torch.int64, device=DEV)
    emb = torch.empty(2, 1, D, device=DEV)
    rows = torch.empty(2, 1, dtype=torch.int64, device=DEV)
    st = L.stream_ptr(torch.device(DEV))
    # a per-row aux kind without aux shards / a replicated-aux kind with aux shards
    rc = lib.rsb_lookup_fwd_sharded_kind(L.KIND_MASK, L.ptr(x), 0, None, 2, 1, D, L.ptr(tp), 1, 8, 8, None, 0, None, None, 0,
                                         None, None, None, L.ptr(emb), None, None, L.ptr(rows), None, None, st)
    assert rc == 10001
    rc = lib.rsb_lookup_fwd_sharded_kind(L.KIND_VANILLA, L.ptr(x), 0, None, 2, 1, D, L.ptr(tp), 1, 8, 8, None, 0, None,
                                         L.ptr(tp), 0, None, None, None, L.ptr(emb), None, None, L.ptr(rows), None, None, st)
    assert rc == 10001
    rc = lib.rsb_lookup_fwd_sharded_kind(L.KIND_VANILLA, L.ptr(x), 0, None, 2, 1, D, None, 1, 8, 8, None, 0, None, None, 0,
                                         None, None, None, L.ptr(emb), None, None, L.ptr(rows), None, None, st)
    assert rc == 10001


# ---------------------------------------------------------------- module level
def _emb_config(name, tmp):
    if name == "qr":
        return {"name": "qr", "divider": 5}
    if name == "qr_add":
        return {"name": "qr", "divider": 7, "operation": "add"}
    if name.startswith("pep_retrain"):
        return {"name": "pep_retrain", "checkpoint_weight_dir": tmp, "sparsity": 0.8}
    if name.startswith("pep"):
        return {"name": "pep", "checkpoint_weight_dir": tmp, "threshold_type": name[4:], "init_threshold": -4.0}
    return {"name": "vanilla"}


def _prepare_retrain_checkpoint(tmp, field_name):
    """The file RetrainPepEmbedding reads (pep_embedding.py:192-203): {dir}/{field_name}/{sparsity}.pth."""
    g = torch.Generator().manual_seed(11)
    n = sum(DIMS)
    os.makedirs(os.path.join(tmp, field_name), exist_ok=True)
    torch.save({"emb.weight": torch.randn(n, D, generator=g) * 0.5, "s": torch.randn(n, D, generator=g) - 1.0},
               os.path.join(tmp, field_name, "0.8.pth"))


def _models(R, S, name, model_kind, tmp, dev, group=None):
    cfg = _emb_config(name, tmp)
    torch.manual_seed(5)
    if model_kind == "deepfm":
        if name.startswith("pep_retrain"):
            _prepare_retrain_checkpoint(tmp, "deepfm")
        full = R.get_ctr_model(DIMS, dict(num_factor=D, hidden_sizes=[32, 16], p_dropout=0.0, use_batchnorm=False,
                                          embedding_config=dict(cfg))).to(dev)
        torch.manual_seed(5)
        sh = S.ShardedDeepFM(DIMS, D, [32, 16], p_dropout=0.0, use_batchnorm=False, embedding_config=dict(cfg),
                             group=group).to(dev)
    else:
        if name.startswith("pep_retrain"):
            _prepare_retrain_checkpoint(tmp, "dcn")
        full = R.get_ctr_model(DIMS, dict(name="dcn_mix", num_factor=D, hidden_sizes=[32], num_layers=2, num_experts=2,
                                          rank=8, p_dropout=0.0, embedding_config=dict(cfg))).to(dev)
        torch.manual_seed(5)
        sh = S.ShardedDCNMix(DIMS, D, [32], num_layers=2, num_experts=2, rank=8, p_dropout=0.0,
                             embedding_config=dict(cfg), group=group).to(dev)
    # same replicated parameters, the big arrays from the single-device module's tensors
    st = {k: v for k, v in full.state_dict().items() if not k.startswith("embedding.")}
    sh.load_state_dict(st, strict=False)
    if name == "vanilla":
        sh.embedding.load_full_weight(full.embedding.get_weight().detach())
    else:
        sh.embedding.load_full_state_dict(full.embedding.state_dict())
    return full, sh


def _full_embedding_state(sh, name):
    if name == "vanilla":
        return {"_emb_module.weight": sh.embedding.gather_full_weight()}
    return sh.embedding.full_state_dict()


def _train(model, x, y, steps, sharded, lr=1e-2):
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr, weight_decay=1e-4)
    crit = torch.nn.BCEWithLogitsLoss()
    outs = []
    for _ in range(steps):
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


MODULE_CASES = [("deepfm", "qr"), ("deepfm", "qr_add"), ("deepfm", "pep_feature_dim"), ("deepfm", "pep_feature"),
                ("deepfm", "pep_dimension"), ("deepfm", "pep_global"), ("deepfm", "pep_retrain"), ("dcn", "vanilla"),
                ("dcn", "qr"), ("dcn", "pep_retrain")]


@pytest.mark.parametrize("model_kind,name", MODULE_CASES)
def test_sharded_variants_world1_match_the_single_device_model(env, tmp_path, model_kind, name):
    RF, L, S = env
    import recsys_benchmark_b200 as R

    dev = torch.device(DEV)
    full, sh = _models(R, S, name, model_kind, str(tmp_path), dev)
    g = torch.Generator().manual_seed(1)
    x = _ids(64, 1).to(dev)
    y = torch.randint(0, 2, (64,), generator=g).float().to(dev)
    a = _train(full, x, y, 3, False)
    b = _train(sh, x, y, 3, True)
    assert torch.equal(a[0], b[0]), "same rows, same kernel arithmetic"
    for s in range(3):
        assert_close(b[s].cpu().numpy(), a[s].cpu().numpy(), what=f"logits step {s}", atol_scale=5e-5)
    ref = full.embedding.state_dict()
    got = _full_embedding_state(sh, name)
    for k, v in ref.items():
        if v.dtype == torch.bool:
            assert torch.equal(got[k], v), k
        else:
            assert_close(got[k].cpu().numpy(), v.detach().cpu().numpy(), what=f"{k} after 3 steps", atol_scale=5e-5)
    if name != "vanilla":
        # the effective table (reference get_weight) straight from the shards
        assert_close(sh.embedding.get_weight().cpu().numpy(), full.embedding.get_weight().detach().cpu().numpy(),
                     what="get_weight", atol_scale=5e-5)


def test_sharded_pep_bookkeeping_matches_the_single_device_plugin(env, tmp_path):
    """get_sparsity / get_num_params / train_callback (scripts/deepfm/train_deepfm_pep.py:63,71,242) over the shards."""
    RF, L, S = env
    import recsys_benchmark_b200 as R

    dev = torch.device(DEV)
    d1, d2 = str(tmp_path / "a"), str(tmp_path / "b")
    cfg = {"name": "pep", "threshold_type": "feature_dim", "init_threshold": -4.0, "sparsity": [0.05, 0.5, 0.999]}
    torch.manual_seed(3)
    full = R.get_embedding(dict(cfg, checkpoint_weight_dir=d1), DIMS, D, field_name="deepfm").to(dev)
    torch.manual_seed(3)
    sh = S.ShardedEmbedding(R.get_embedding(dict(cfg, checkpoint_weight_dir=d2), DIMS, D, field_name="deepfm"))
    sh.load_full_state_dict(full.state_dict())
    assert sh.get_num_params() == full.get_num_params()
    assert sh.get_sparsity() == full.get_sparsity()
    assert sh.get_sparsity(True) == full.get_sparsity(True)
    assert sh.sparsity == full.sparsity and sh.threshold_type == "feature_dim"          # read-through attributes
    full.train_callback()
    sh.train_callback()
    assert sh._cur_min_spar_idx == full._cur_min_spar_idx >= 1
    for target in full.sparsity[: full._cur_min_spar_idx]:
        a = torch.load(os.path.join(full.checkpoint_weight_dir, f"{target}.pth"), map_location="cpu")
        b = torch.load(os.path.join(sh.checkpoint_weight_dir, f"{target}.pth"), map_location="cpu")
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), k
    # QR: parameter count of the FULL tables
    qr_full = R.get_embedding({"name": "qr", "divider": 5}, DIMS, D).to(dev)
    qr_sh = S.ShardedEmbedding(R.get_embedding({"name": "qr", "divider": 5}, DIMS, D))
    assert qr_sh.get_num_params() == qr_full.get_num_params()


def _worker2(rank, world, port, ret, model_kind, name, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import recsys_benchmark_b200 as R
        from recsys_benchmark_b200 import sharded as S

        full, sh = _models(R, S, name, model_kind, os.path.join(tmp, str(rank)), dev)
        # BatchNorm statistics are per rank (as with DDP): with batch statistics the global-batch run is not the
        # reference for a rank's slice, so the DCN tail's BatchNorm layers run on their (fixed) running statistics here
        for m in list(full.modules()) + list(sh.modules()):
            if isinstance(m, torch.nn.BatchNorm1d):
                m.eval()
        g = torch.Generator().manual_seed(2)
        x = _ids(128, 2)
        y = torch.randint(0, 2, (128,), generator=g).float()
        a = _train(full, x.to(dev), y.to(dev), 3, False)          # every rank: the single-GPU run on the global batch
        b = _train(sh, x[rank::world].to(dev), y[rank::world].to(dev), 3, True)
        assert torch.equal(a[0][rank::world], b[0]), "forward over peer shards differs from the single-GPU gather"
        for s in range(3):
            assert_close(b[s].cpu().numpy(), a[s][rank::world].cpu().numpy(), what=f"logits step {s}", atol_scale=1e-4)
        ref = full.embedding.state_dict()
        got = _full_embedding_state(sh, name)
        for k, v in ref.items():
            if v.dtype == torch.bool:
                assert torch.equal(got[k], v), k
            else:
                assert_close(got[k].cpu().numpy(), v.detach().cpu().numpy(), what=f"{k}", atol_scale=1e-4)
        if name.startswith("pep_") and not name.startswith("pep_retrain"):
            # the plugin's bookkeeping is global over the shards (counts all-reduced); the two tables agree to rounding,
            # so the counts of surviving weights may differ by the few elements that sit on a threshold
            n_sh, n_full = sh.embedding.get_num_params(), full.embedding.get_num_params()
            assert abs(n_sh - n_full) <= 8, (n_sh, n_full)
        sh.embedding.shards.close()          # collective: unmap the peers, barrier, free
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("model_kind,name", [("deepfm", "qr"), ("deepfm", "pep_feature_dim"), ("deepfm", "pep_retrain"),
                                             ("dcn", "vanilla")])
def test_sharded_variants_world2_match_single_gpu(tmp_path, model_kind, name):
    import __graft_entry__ as G

    G.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29300 + os.getpid() % 500 + MODULE_CASES.index((model_kind, name))
    mp.spawn(_worker2, args=(2, port, ret, model_kind, name, str(tmp_path)), nprocs=2, join=True)
    assert dict(ret) == {0: 1, 1: 1}
