"""GPU parity tests through the public (reference-shaped) API: our DeepFM / DCN_Mix and
embedding plugins loaded with the reference's own state dicts (tests/golden/*.npz, made by
the unmodified reference) must reproduce its logits, gradients and post-step weights.
Plus edge cases (empty / ragged inputs, int32 ids, out-of-range ids, bag modes) and
size-independent properties at the BASELINE.json sizes."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import assert_close, load_golden, sub
from tests.test_host_api import CASES, build_from_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CRITEO_DIMS = [49, 101, 126, 45, 223, 118, 84, 76, 95, 9, 30, 40, 75, 1458, 555, 193949, 138801, 306, 19, 11970, 634,
               4, 42646, 5178, 192773, 3175, 27, 11422, 181075, 11, 4654, 2032, 5, 189657, 18, 16, 59697, 86, 45571]


@pytest.fixture(scope="module")
def R():
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200 as r

    return r


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def _grad_np(p):
    g = p.grad
    return (g.to_dense() if g.is_sparse else g).detach().cpu().numpy()


def _bias_before_batchnorm(model):
    out = set()
    for seq_name in ("_deep_branch", "_dnn", "embedding._seq"):
        seq = model
        for part in seq_name.split("."):
            seq = getattr(seq, part, None)
            if seq is None:
                break
        if seq is None:
            continue
        mods = list(seq)
        for i, m in enumerate(mods[:-1]):
            if isinstance(m, torch.nn.Linear) and isinstance(mods[i + 1], torch.nn.BatchNorm1d):
                out.add(f"{seq_name}.{i}.bias")
    return out


def _run_steps(R, name, emb_cfg, opt_cfg, steps, tmp_path=None, pre_step=None, atol_scale=1e-5, state_extra=None,
               after_keys=()):
    g, model, state = build_from_golden(name, emb_cfg)
    if state_extra is not None:
        state.update(state_extra(g))
    model.load_state_dict(state, strict=True)
    model.to(DEV)
    # eval logits
    model.eval()
    if pre_step is not None:
        pre_step(model, "eval", g)
    with torch.no_grad():
        logits = model(_t(g["x"]))
    assert_close(logits.cpu().numpy(), g["logits_eval"], what=f"{name} eval logits", atol_scale=atol_scale)
    opts = R.get_optimizers(model, opt_cfg) if opt_cfg else []
    crit = torch.nn.BCEWithLogitsLoss()
    model.train()
    for s in range(steps):
        if pre_step is not None:
            pre_step(model, f"step{s}", g)
        logits = model(_t(g[f"step{s}/x"]))
        loss = crit(logits, _t(g[f"step{s}/y"]).float())
        for o in opts:
            o.zero_grad()
        for p in model.parameters():
            p.grad = None
        loss.backward()
        assert_close(logits.detach().cpu().numpy(), g[f"step{s}/logits"], what=f"{name} step{s} logits",
                     atol_scale=atol_scale)
        noise_keys = _bias_before_batchnorm(model)
        for k, p in model.named_parameters():
            key = f"step{s}/grad/{k}"
            if k in noise_keys:
                # d loss / d (Linear bias feeding BatchNorm) is zero by construction: both sides hold fp32 noise
                assert float(p.grad.abs().max()) < 1e-5 and float(np.abs(g[key]).max()) < 1e-5
                continue
            if key in g:
                assert p.grad is not None, f"{k} has no grad"
                # atol floor: gradients that are zero by construction (a Linear bias feeding BatchNorm) are fp32 noise ~1e-8
                assert_close(_grad_np(p), g[key], what=f"{name} step{s} grad {k}", atol_scale=2e-5, atol_floor=2e-7)
            else:
                assert p.grad is None or float(p.grad.abs().sum()) == 0.0 or not p.requires_grad, k
        for o in opts:
            o.step()
        if opts:
            after = sub(g, f"step{s}/after/")
            cur = model.state_dict()
            for k in ["embedding.p_weight", "embedding.q_weight", "embedding.p_threshold", "embedding.q_threshold",
                      "embedding._emb_module.weight", "embedding.emb1.weight", "embedding.emb2.weight",
                      "embedding.emb.weight", "embedding.s", "embedding._weight",
                      "embedding._mask_e_module._t_param", "fc.weight", "_bias", *after_keys]:
                if k in after:
                    # floor: Adam's g / (sqrt(v) + eps) turns a rounding-level difference of a near-zero gradient
                    # element into a visible fraction of lr
                    lr = (opt_cfg or {}).get("learning_rate", 0.0)
                    assert_close(cur[k].cpu().numpy(), after[k], what=f"{name} step{s} after {k}", atol_scale=5e-5,
                                 atol_floor=max(2.5e-3 * lr, 1e-30))
    return g, model


ADAM = dict(learning_rate=1e-2, weight_decay=1e-4)
SPARSE_ADAM = dict(learning_rate=1e-2, weight_decay=1e-4, sparse=True)
SPARSE_SGD = dict(learning_rate=1e-1, weight_decay=1e-4, sparse=True, optimizer="sgd")


def test_deepfm_vanilla_dense_adam(R):
    _run_steps(R, "deepfm_vanilla_adam", CASES["deepfm_vanilla_adam"], ADAM, 2)


def test_deepfm_vanilla_int32_ids(R):
    g, model = _run_steps(R, "deepfm_vanilla_int32", {"name": "vanilla"}, None, 1)
    assert g["x"].dtype == np.int32


def test_deepfm_vanilla_sparse_coo_grad_and_sparse_adam(R):
    g, model = _run_steps(R, "deepfm_vanilla_sparse_adam", CASES["deepfm_vanilla_sparse_adam"], SPARSE_ADAM, 3)
    gr = model.embedding._emb_module.weight.grad
    b, f = g["step2/x"].shape
    assert gr.is_sparse and not gr.is_coalesced() and gr._nnz() == b * f   # same layout as nn.Embedding(sparse=True)
    assert gr._indices().dtype == torch.int64 and tuple(gr._indices().shape) == (1, b * f)


@pytest.mark.parametrize("name,cfg", [("deepfm_vanilla_sparse_adam", dict(SPARSE_ADAM, fused_sparse=True)),
                                      ("deepfm_vanilla_sparse_sgd", dict(SPARSE_SGD, fused_sparse=True))])
def test_deepfm_fused_sparse_update_matches_reference_steps(R, name, cfg):
    g, model, state = build_from_golden(name, {"name": "vanilla", "sparse": True})
    model.load_state_dict(state, strict=True)
    model.to(DEV).train()
    opts = R.get_optimizers(model, cfg)
    crit = torch.nn.BCEWithLogitsLoss()
    steps = 3 if "adam" in name else 2
    for s in range(steps):
        logits = model(_t(g[f"step{s}/x"]))
        loss = crit(logits, _t(g[f"step{s}/y"]).float())
        for o in opts:
            o.zero_grad()
        loss.backward()
        assert model.embedding._emb_module.weight.grad is None   # no gradient tensor is materialised
        for o in opts:
            o.step()
        assert_close(model.embedding._emb_module.weight.detach().cpu().numpy(),
                     g[f"step{s}/after/embedding._emb_module.weight"], what=f"fused step {s}", atol_scale=5e-5)


def test_deepfm_vanilla_sparse_sgd(R):
    _run_steps(R, "deepfm_vanilla_sparse_sgd", {"name": "vanilla", "sparse": True}, SPARSE_SGD, 2)


@pytest.mark.parametrize("op", ["mult", "add", "cat"])
def test_deepfm_qr(R, op):
    _run_steps(R, f"deepfm_qr_{op}", CASES[f"deepfm_qr_{op}"], ADAM, 2)


@pytest.mark.parametrize("tt", ["feature_dim", "feature", "dimension", "global"])
def test_deepfm_pep(R, tt, tmp_path):
    _run_steps(R, f"deepfm_pep_{tt}", {"name": "pep", "checkpoint_weight_dir": str(tmp_path), "threshold_type": tt},
               ADAM, 1)


@pytest.mark.parametrize("sparse", [False, True])
def test_deepfm_pep_retrain(R, sparse, tmp_path):
    ck = load_golden("pep_retrain_ckpt")
    (tmp_path / "deepfm").mkdir()
    torch.save({"emb.weight": torch.from_numpy(ck["weight"]), "s": torch.from_numpy(ck["s"])},
               tmp_path / "deepfm" / "0.5.pth")
    name = "deepfm_pep_retrain" + ("_sparse" if sparse else "")
    g, model = _run_steps(R, name, {"name": "pep_retrain", "checkpoint_weight_dir": str(tmp_path), "sparsity": 0.5,
                                    "sparse": sparse}, SPARSE_ADAM if sparse else ADAM, 1)
    assert model.embedding.emb.weight.grad.is_sparse == sparse


def test_deepfm_cerp_retrain(R, tmp_path):
    from tests.test_host_api import make_cerp_retrain_dir

    g, model = _run_steps(R, "deepfm_cerp_retrain", make_cerp_retrain_dir(tmp_path), ADAM, 2)
    emb = model.embedding
    assert emb.get_num_params() == int(g["state/embedding.q_mask"].sum() + g["state/embedding.p_mask"].sum())
    assert tuple(emb.get_weight().shape) == (int(g["field_dims"].sum()), 8)


def test_deepfm_cerp(R):
    g, model = _run_steps(R, "deepfm_cerp", CASES["deepfm_cerp"], ADAM, 2)
    emb = model.embedding
    # bookkeeping the CERP trainer calls (src/trainer/deepfm.py:142-248)
    sp, nnz = emb.get_sparsity(True)
    st = {k: v.detach().cpu().numpy() for k, v in emb.state_dict().items()}
    ref_nnz = int(np.count_nonzero(O.pep_soft_threshold(st["p_weight"], st["p_threshold"])) +
                  np.count_nonzero(O.pep_soft_threshold(st["q_weight"], st["q_threshold"])))
    assert abs(nnz - ref_nnz) <= 1
    emb.apply_pruning()   # get_prune_loss reads the tables pruned by the last forward (cerp_embedding.py:205-207)
    loss = emb.get_prune_loss()
    ref = O.cerp_prune_loss(st["p_weight"].astype(np.float64), st["q_weight"].astype(np.float64),
                            st["p_threshold"].astype(np.float64), st["q_threshold"].astype(np.float64))
    assert abs(float(loss) - ref) < 1e-3 * abs(ref) + 1e-6
    assert tuple(emb.get_weight().shape) == (int(g["field_dims"].sum()), 8)


def _optembed_pre(model, tag, g):
    if tag == "eval":
        model.embedding.get_weight()
        return
    # replay the reference's mask-D draw: we make the same torch.randint call, so seeding the
    # CPU generator is not enough on CUDA -> inject the recorded draw instead
    k = _t(g[f"{tag}/mask_d_idx"])
    model.embedding._mask_d_for = lambda b, f, device, _k=k: _k


@pytest.mark.parametrize("name", ["deepfm_optembed", "deepfm_optembed_l2", "deepfm_optembed_d"])
def test_deepfm_optembed(R, name):
    steps = 2 if name == "deepfm_optembed" else 1
    g, model = _run_steps(R, name, CASES[name], ADAM if name == "deepfm_optembed" else None, steps,
                          pre_step=_optembed_pre)


def test_optembed_draws_mask_with_the_reference_torch_call(R):
    emb = R.get_embedding({"name": "deepfm_optembed"}, [5, 4, 6], 8).to(DEV).train()
    torch.manual_seed(11)
    k1 = emb._mask_d_for(7, 3, torch.device(DEV))
    torch.manual_seed(11)
    k2 = torch.randint(0, 8, size=(7, 3), device=DEV)
    assert torch.equal(k1, k2) and k1.dtype == torch.int64


def test_optembed_fresh_eval_raises_like_reference(R):
    emb = R.get_embedding({"name": "deepfm_optembed"}, [5, 4, 6], 8).to(DEV).eval()
    with pytest.raises(RuntimeError, match="2-D"):
        emb(torch.zeros(2, 3, dtype=torch.long, device=DEV))


def test_dcn_mix_matches_reference(R):
    g = load_golden("dcn_mix")
    fd = [int(v) for v in g["field_dims"]]
    cfg = dict(name="dcn_mix", num_factor=4, hidden_sizes=[12], num_layers=2, num_experts=3, rank=5, p_dropout=0.0,
               compile_model=False, embedding_config={"name": "vanilla"})
    model = R.get_ctr_model(fd, cfg)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "state/").items()}, strict=True)
    model.to(DEV).eval()
    with torch.no_grad():
        assert_close(model(_t(g["x"])).cpu().numpy(), g["logits_eval"], what="dcn eval logits")
    model.train()
    logits = model(_t(g["x"]))
    torch.nn.BCEWithLogitsLoss()(logits, _t(g["y"]).float()).backward()
    assert_close(logits.detach().cpu().numpy(), g["logits"], what="dcn logits")
    noise_keys = _bias_before_batchnorm(model)
    for k, p in model.named_parameters():
        if f"grad/{k}" in g and k not in noise_keys:
            assert_close(_grad_np(p), g[f"grad/{k}"], what=f"dcn grad {k}", atol_scale=3e-5, atol_floor=2e-7)


# ------------------------------------------------------------- plugin API ---
def test_embedding_api_goldens(R, tmp_path):
    a = load_golden("embedding_api")
    fd = [int(v) for v in a["field_dims"]]
    ids1, ids2 = _t(a["ids1"]), _t(a["ids2"])
    for op in ["mult", "add", "cat"]:
        emb = R.get_embedding({"name": "qr", "divider": 4, "operation": op}, fd, 6)
        emb.load_state_dict({k[len(f"qr_{op}/state/"):]: torch.from_numpy(v) for k, v in a.items()
                             if k.startswith(f"qr_{op}/state/")})
        emb.to(DEV)
        np.testing.assert_array_equal(emb(ids1).detach().cpu().numpy(), a[f"qr_{op}/out1"])
        np.testing.assert_array_equal(emb(ids2).detach().cpu().numpy(), a[f"qr_{op}/out2"])
        np.testing.assert_array_equal(emb.get_weight().detach().cpu().numpy(), a[f"qr_{op}/get_weight"])
    # OptEmbed retrain
    emb = R.get_embedding({"name": "deepfm_optembed_retrain"}, fd, 6)
    with torch.no_grad():
        emb._weight.copy_(torch.from_numpy(a["optretrain/weight"]))
    m = emb.init_mask(mask_e=torch.from_numpy(a["optretrain/mask_e"]), mask_d=torch.from_numpy(a["optretrain/mask_d"]))
    np.testing.assert_array_equal(m.cpu().numpy(), a["optretrain/mask"])
    emb.to(DEV)
    np.testing.assert_array_equal(emb(ids2).detach().cpu().numpy(), a["optretrain/out2"])
    sp, nnz = emb.get_sparsity(True)
    assert nnz == int(a["optretrain/nnz"]) and abs(sp - float(a["optretrain/sparsity"])) < 1e-12
    # OptEmbed eval weight, explicit mask_d, bookkeeping
    emb = R.get_embedding({"name": "deepfm_optembed"}, fd, 6)
    with torch.no_grad():
        emb._weight.copy_(torch.from_numpy(a["opteval/weight"]))
        emb._mask_e_module._t_param.copy_(torch.from_numpy(a["opteval/t"]))
    emb.to(DEV).eval()
    np.testing.assert_array_equal(emb.get_weight().cpu().numpy(), a["opteval/w_plain"])
    np.testing.assert_array_equal(emb.get_weight(torch.tensor([1, 4, 2])).cpu().numpy(), a["opteval/w_maskd"])
    assert_close(emb.get_l_s().detach().cpu().numpy(), a["opteval/l_s"], what="l_s")
    sp, nnz = emb.get_sparsity(True)
    assert nnz == int(a["opteval/nnz"])
    np.testing.assert_array_equal(emb.get_mask_e().numpy(), a["opteval/mask_e"])
    np.testing.assert_array_equal(emb.get_submask().numpy(), a["opteval/submask"])
    # PEP bookkeeping
    emb = R.get_embedding({"name": "pep", "checkpoint_weight_dir": str(tmp_path)}, fd, 6, field_name="deepfm")
    with torch.no_grad():
        emb.emb.weight.copy_(torch.from_numpy(a["pep/weight"]))
        emb.s.copy_(torch.from_numpy(a["pep/s"]))
    emb.to(DEV)
    sp, nnz = emb.get_sparsity(True)
    assert nnz == int(a["pep/nnz"])
    assert_close(emb.get_weight().cpu().numpy(), a["pep/get_weight"], what="pep get_weight")
    assert_close(emb(ids1).detach().cpu().numpy(), a["pep/out1"], what="pep 1-D")
    emb.sparsity = [0.1, 0.99]
    emb.train_callback()
    assert (tmp_path / "deepfm" / "0.1.pth").exists() and emb._cur_min_spar_idx == 1


@pytest.mark.parametrize("name", ["vanilla", "qr", "deepfm_optembed_d"])
def test_plugin_shapes_like_reference_tests(R, name):
    """tests/test_emb.py:124-164 of the reference: [11,3] ids -> (11,3,7); 1-D ids -> (11,7)."""
    emb = R.get_embedding({"name": name}, [5, 6, 7], 7).to(DEV)
    x = torch.randint(0, 18, (11, 3), device=DEV)
    assert tuple(emb(x).shape) == (11, 3, 7)
    if name != "deepfm_optembed_d":
        assert tuple(emb(x[:, 0]).shape) == (11, 7)
    assert tuple(emb.get_weight().shape) == (18, 7)


@pytest.mark.parametrize("mode", ["sum", "mean", "max"])
def test_bag_modes(R, mode):
    emb = R.get_embedding({"name": "vanilla"}, [50], 16, mode=mode).to(DEV)
    x = torch.randint(0, 50, (9, 4), device=DEV)
    ref = torch.nn.functional.embedding(x, emb.get_weight())
    ref = {"sum": ref.sum(1), "mean": ref.mean(1), "max": ref.max(1).values}[mode]
    torch.testing.assert_close(emb(x), ref, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------- edge cases ---
def test_empty_batch(R):
    model = R.get_ctr_model([5, 6, 7], dict(num_factor=16, hidden_sizes=[8], p_dropout=0.0)).to(DEV).eval()
    out = model(torch.zeros(0, 3, dtype=torch.long, device=DEV))
    assert tuple(out.shape) == (0,)


def test_out_of_range_ids_raise_index_error_when_validating(R):
    emb = R.get_embedding({"name": "vanilla"}, [5, 6], 16).to(DEV)
    emb.validate = True
    with pytest.raises(IndexError):
        emb(torch.tensor([[0, 11]], device=DEV))
    with pytest.raises(IndexError):
        emb(torch.tensor([[-1, 3]], device=DEV))
    assert tuple(emb(torch.tensor([[4, 10]], device=DEV)).shape) == (1, 2, 16)   # max valid ids


def test_out_of_range_ids_surface_at_the_next_train_eval_switch_without_validate(R):
    """Default mode (no per-call sync): the gather flags the bad id, and the next model.eval() / model.train() - the
    epoch boundaries of the reference trainer - raises the IndexError F.embedding would have raised."""
    model = R.get_ctr_model([5, 6, 7], dict(num_factor=16, hidden_sizes=[8], p_dropout=0.0)).to(DEV).train()
    model(torch.tensor([[0, 1, 2]], device=DEV))
    model.eval()                                            # clean so far
    model(torch.tensor([[0, 6, 2]], device=DEV))            # field 1 has 6 ids: 6 is out of range... for the table it
    model.train()                                           # is still a valid global row, like in the reference
    model(torch.tensor([[0, 1, 7 + 11]], device=DEV))       # global row 29 >= 18: out of range
    with pytest.raises(IndexError):
        model.eval()
    model.train()                                           # the flag was cleared by the raise


def test_fused_sparse_adam_accumulates_backwards_like_torch_sparse_adam(R):
    """Two backward passes before one optimizer step = ONE SparseAdam step on the coalesced sum."""
    torch.manual_seed(1)
    dims = [50, 7, 300, 11]
    ours = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[32], p_dropout=0.0,
                                      embedding_config={"name": "vanilla", "sparse": True})).to(DEV).train()
    ref = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[32], p_dropout=0.0,
                                     embedding_config={"name": "vanilla", "sparse": True})).to(DEV).train()
    ref.load_state_dict(ours.state_dict())
    o_ours = R.FusedSparseAdam(ours.embedding, lr=1e-2)
    o_ref = torch.optim.SparseAdam(list(ref.embedding.parameters()), lr=1e-2)
    g = torch.Generator().manual_seed(2)
    for step in range(2):
        o_ours.zero_grad()
        o_ref.zero_grad()
        for micro in range(2):
            x = torch.stack([torch.randint(0, d, (64,), generator=g) for d in dims], 1).to(DEV)
            y = torch.randint(0, 2, (64,), generator=g).float().to(DEV)
            for m in (ours, ref):
                torch.nn.functional.binary_cross_entropy_with_logits(m(x), y).backward()
        o_ours.step()
        o_ref.step()
        assert o_ours.state[ours.embedding.get_weight()]["step"] == step + 1
        assert_close(ours.embedding.get_weight().detach().cpu().numpy(),
                     ref.embedding.get_weight().detach().cpu().numpy(), what=f"step {step}", atol_scale=2e-5)
    o_ours.detach_from_module()


def test_unsupported_row_width_fails_loudly(R):
    emb = R.get_embedding({"name": "vanilla"}, [5, 6], 50).to(DEV)   # 50 % 4 != 0 and > 32
    with pytest.raises(RuntimeError, match="unsupported row width"):
        emb(torch.tensor([[0, 5]], device=DEV))


@pytest.mark.parametrize("d", [7, 12, 16, 64])
def test_odd_widths_match_torch(R, d):
    model = R.get_ctr_model([50, 60, 70], dict(num_factor=d, hidden_sizes=[8], p_dropout=0.0)).to(DEV)
    x = torch.stack([torch.randint(0, n, (33,), device=DEV) for n in [50, 60, 70]], 1)
    logits = model(x)
    w = model.embedding.get_weight()
    rows = x + model.offsets
    e = torch.nn.functional.embedding(rows, w)
    y = model.fc(rows) + model._bias + 0.5 * (e.sum(1).pow(2) - e.pow(2).sum(1)).sum(1, keepdim=True)
    ref = (y + model._deep_branch(e.reshape(33, -1))).squeeze(-1)
    torch.testing.assert_close(logits, ref, rtol=1e-5, atol=1e-5)
    g1 = torch.autograd.grad(logits.sum(), [w, model.fc.weight])
    g2 = torch.autograd.grad(ref.sum(), [w, model.fc.weight])
    torch.testing.assert_close(g1[0], g2[0], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(g1[1], g2[1], rtol=1e-4, atol=1e-5)


# ----------------------------------------- properties at BASELINE.json sizes ---
@pytest.mark.parametrize("emb_cfg,b", [({"name": "vanilla"}, 2048), ({"name": "qr", "divider": 5}, 2048),
                                       ({"name": "vanilla"}, 16384), ({"name": "qr", "divider": 20}, 2048),
                                       ({"name": "qr"}, 4096), ({"name": "qr", "divider": 2, "operation": "add"}, 2048),
                                       ({"name": "qr", "divider": 5}, 65536)])
def test_criteo_shape_properties(R, emb_cfg, b):
    torch.manual_seed(3)
    model = R.get_ctr_model(CRITEO_DIMS, dict(num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=0.0,
                                              use_batchnorm=False, embedding_config=dict(emb_cfg))).to(DEV)
    gen = torch.Generator().manual_seed(2023)
    x = torch.stack([torch.randint(0, d, (b,), generator=gen) for d in CRITEO_DIMS], 1).to(DEV)
    # (1) fused forward == unfused composition of the plugin forward and torch FM ops
    logits = model(x)
    rows = x + model.offsets
    e = model.embedding(rows)
    y = model.fc(rows) + model._bias + 0.5 * (e.sum(1).pow(2) - e.pow(2).sum(1)).sum(1, keepdim=True)
    ref = (y + model._deep_branch(e.reshape(b, -1))).squeeze(-1)
    assert_close(logits.detach().cpu().numpy(), ref.detach().cpu().numpy(), what="fused vs composed", atol_scale=2e-5)
    # (2) int32 ids give bit-identical logits
    assert torch.equal(model(x.int()), logits)
    # (2b) the gather reports max |emb| through its per-warp slots (the dense tail's operand scale comes from it)
    import recsys_benchmark_b200.functional as RF_
    emb_l, _ = model.embedding.lookup(x, model.offsets, model.fc.weight, model._bias)
    slots = RF_.amax_slots_of(emb_l)
    assert slots is not None and float(slots.max()) == float(emb_l.abs().max())
    with torch.no_grad():
        assert RF_.amax_slots_of(model.embedding.lookup(x, model.offsets, model.fc.weight, model._bias)[0]) is None
    # (3) gradients: fused backward == composed backward; checksum of checksums
    params = [p for p in model.embedding.parameters()] + [model.fc.weight]
    g1 = torch.autograd.grad(logits.square().sum(), params)
    g2 = torch.autograd.grad(ref.square().sum(), params)
    for a_, b_ in zip(g1, g2):
        assert_close(a_.cpu().numpy(), b_.cpu().numpy(), what="grad fused vs composed", atol_scale=5e-5)
    # (4) determinism of the sorted segmented scatter-add (the big table: vanilla weight / QR emb2;
    #     QR emb1 accumulates in shared memory with float atomics and fc with global atomics)
    g3 = torch.autograd.grad(model(x).square().sum(), params)
    big = max(range(len(params) - 1), key=lambda i: params[i].numel())
    assert torch.equal(g1[big], g3[big])
    if emb_cfg["name"] == "vanilla":
        # the whole vanilla backward is order-fixed: fc.weight.grad comes from the sorted lookups too
        assert torch.equal(g1[-1], g3[-1])
    # (5) untouched rows have exactly zero gradient
    if emb_cfg["name"] == "vanilla":
        touched = torch.zeros(sum(CRITEO_DIMS), dtype=torch.bool, device=DEV)
        touched[rows.reshape(-1)] = True
        assert float(g1[0][~touched].abs().sum()) == 0.0


def test_device_prefetcher_yields_the_same_batches_in_order(R):
    from recsys_benchmark_b200.data import DevicePrefetcher

    batches = [(torch.randint(0, 100, (64, 5), dtype=torch.int32), torch.rand(64)) for _ in range(7)]
    n = 0
    for (x, y), (xd, yd) in zip(batches, DevicePrefetcher(iter(batches), DEV)):
        # double-buffered: a yielded batch stays valid until the batch after the next one is staged
        assert xd.is_cuda and yd.is_cuda
        assert torch.equal(xd.cpu(), x) and torch.equal(yd.cpu(), y)
        n += 1
    assert n == 7
    assert list(DevicePrefetcher(iter([]), DEV)) == []


# ------------------------------------------------------- pruned CSR inference path (f-4) ---
def _pruned_model(R, g, compact=False):
    fd = [int(v) for v in g["field_dims"]]
    st = sub(g, "state/")
    model = R.get_ctr_model(fd, dict(num_factor=8, hidden_sizes=[16, 8], p_dropout=0.1, use_batchnorm=True,
                                     embedding_config={"name": "vanilla"}))
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)
    model.embedding = R.PrunedEmbedding.from_other_emb(model.embedding, compact=compact)
    return model.to(DEV).eval(), fd


@pytest.mark.parametrize("compact", [False, True])
def test_pruned_csr_matches_reference(R, compact):
    g = load_golden("pruned_csr")
    model, fd = _pruned_model(R, g, compact)
    emb = model.embedding
    assert emb.is_cuda
    emb.to_cuda()
    rows = _t(O.add_offsets(g["x"], O.field_offsets(g["field_dims"])))
    with torch.no_grad():
        np.testing.assert_array_equal(emb(rows).cpu().numpy(), g["emb"])                   # bit exact
        np.testing.assert_array_equal(emb(rows[:, 2].contiguous()).cpu().numpy(), g["emb_1d"])
        np.testing.assert_array_equal(emb(rows.reshape(2, -1, rows.shape[1])).cpu().numpy(),
                                      g["emb"].reshape(2, -1, *g["emb"].shape[1:]))
        np.testing.assert_array_equal(emb.get_weight().cpu().numpy(), g["weight_dense"])
        assert_close(model(_t(g["x"])).cpu().numpy(), g["logits"], what="logits")
        assert_close(model(_t(g["x_int32"])).cpu().numpy(), g["logits_int32"], what="logits int32")
        assert model(_t(g["x"])[:0]).shape == (0,)                                         # empty batch


@pytest.mark.parametrize("d,keep", [(16, 0.2), (16, 1.0), (16, 0.0), (32, 0.5), (5, 0.5)])
def test_pruned_csr_equals_dense_gather_at_criteo_shape(R, d, keep):
    torch.manual_seed(3)
    n = sum(CRITEO_DIMS)
    w = torch.randn(n, d, device=DEV) * (torch.rand(n, d, device=DEV) < keep)
    emb = R.PrunedEmbedding.from_weight(w, compact=True)
    van = R.PrunedEmbedding.from_weight(w)
    x = torch.stack([torch.randint(0, v, (4096,)) for v in CRITEO_DIMS], 1).to(DEV)
    offsets = torch.tensor([0] + CRITEO_DIMS[:-1]).cumsum(0).to(DEV)
    fc = torch.randn(n, 1, device=DEV)
    bias = torch.randn(1, device=DEV)
    with torch.no_grad():
        e1, y1 = emb.lookup(x.int(), offsets, fc, bias)
        e2, y2 = van.lookup(x, offsets, fc, bias)
    ref = w[x + offsets]
    assert torch.equal(e1, ref) and torch.equal(e2, ref)
    assert torch.equal(y1, y2)
    yref = (fc[x + offsets].sum(1) + bias)[:, 0] + 0.5 * (ref.sum(1).pow(2) - ref.pow(2).sum(1)).sum(1)
    assert_close(y1.cpu().numpy(), yref.cpu().numpy(), what="y_fm")


def test_pruned_csr_out_of_range_and_wide_rows(R):
    w = torch.randn(10, 8, device=DEV)
    emb = R.PrunedEmbedding.from_weight(w)
    emb.validate = True
    with pytest.raises(IndexError):
        emb(torch.tensor([[3, 10]], device=DEV))
    with pytest.raises(RuntimeError):                                   # D > 32 is not supported by this kernel
        R.PrunedEmbedding.from_weight(torch.randn(10, 64, device=DEV))(torch.tensor([[1]], device=DEV))


# ------------------------------------------------------- deep hash embedding (f-3) ---
DHE_CASES = {
    "deepfm_dhe": {"name": "dhe", "inp_size": 32, "hidden_sizes": [16]},
    "deepfm_dhe_v2": {"name": "dhe", "inp_size": 24, "hidden_sizes": [], "use_bn": 1, "compute_v2": True},
    "deepfm_dhe_nobn": {"name": "dhe", "inp_size": 32, "hidden_sizes": [12, 16], "use_bn": 0},
}


def _dhe_extra(g):
    return {"embedding._extra_state": {"_prefix": int(g["dhe/prefix"])}}


@pytest.mark.parametrize("name", sorted(DHE_CASES))
def test_dhe_codes_are_bit_exact(R, name):
    g = load_golden(name)
    fd = [int(v) for v in g["field_dims"]]
    emb = R.get_embedding(DHE_CASES[name], fd, 8)
    emb.load_state_dict({**{k[len("embedding."):]: torch.from_numpy(np.asarray(v)) for k, v in sub(g, "state/").items()
                            if k.startswith("embedding.")}, "_extra_state": {"_prefix": int(g["dhe/prefix"])}})
    emb.to(DEV)
    np.testing.assert_array_equal(emb._cache.cpu().numpy(), g["dhe/cache"])            # fast modulo path
    np.testing.assert_array_equal(emb.encode(_t(g["dhe/ids"])).cpu().numpy(), g["dhe/hash_batch"])  # generic int64
    ids = torch.arange(len(g["dhe/cache"]), device=DEV)
    np.testing.assert_array_equal(emb.encode(ids).cpu().numpy(), g["dhe/cache"])       # both paths agree
    np.testing.assert_array_equal(emb.encode(ids.int(), in_table=True).cpu().numpy(), g["dhe/cache"])


@pytest.mark.parametrize("name", sorted(DHE_CASES))
def test_deepfm_dhe_matches_reference(R, name):
    g = load_golden(name)
    keys = [k[len("step0/after/"):] for k in g if k.startswith("step0/after/embedding._seq") and
            (k.endswith("weight") or k.endswith("bias"))]
    if DHE_CASES[name].get("use_bn", 2) == 2:
        # a Linear bias feeding BatchNorm has a zero gradient by construction; Adam normalises the fp32 noise
        keys = [k for k in keys if not (k.endswith(".bias") and g["state/" + k[:-4] + "weight"].ndim == 2)]
    _run_steps(R, name, DHE_CASES[name], ADAM, 2, state_extra=_dhe_extra, after_keys=keys)


def test_dhe_codes_at_scale_match_oracle(R):
    """k = 1024 codes at the Criteo table size (prefix pushes ids past 2^20): bit-exact against the int64 oracle."""
    R.DHEmbedding.COUNTER = 123456
    emb = R.DHEmbedding(CRITEO_DIMS, 16, None, 1024, []).to(DEV)
    ids = torch.randint(0, sum(CRITEO_DIMS), (4096,), device=DEV)
    ref = O.dhe_universal_hash(ids.cpu().numpy(), emb._prefix, emb._slopes.cpu().numpy(), emb._bias.cpu().numpy(),
                               emb._primes_choices.cpu().numpy())
    np.testing.assert_array_equal(emb.encode(ids, in_table=True).cpu().numpy(), ref)
    np.testing.assert_array_equal(emb.encode(ids).cpu().numpy(), ref)
    assert float(ref.min()) >= -1.0 and float(ref.max()) <= 1.0


def test_deferred_scalar_returns_each_value_one_push_late(R):
    from recsys_benchmark_b200.data import DeferredScalar

    r = DeferredScalar(DEV)
    vals = [torch.tensor(float(i) * 1.5, device=DEV) for i in range(5)]
    got = [r.push(v) for v in vals]
    assert got == [None, 0.0, 1.5, 3.0, 4.5] and r.flush() == 6.0


def test_graphed_train_step_matches_eager_steps(R):
    """The whole step replayed from one CUDA graph produces the same losses / weights as eager launches."""
    from recsys_benchmark_b200.graphed import GraphedTrainStep

    dims = [50, 7, 300, 12, 1000, 4]

    def build():
        torch.manual_seed(11)
        m = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[64, 32], p_dropout=0.0, use_batchnorm=False,
                                       embedding_config={"name": "qr", "divider": 5})).to(DEV).train()
        return m, R.get_optimizers(m, dict(learning_rate=1e-2, weight_decay=1e-5, fused_adam=True, capturable=True))

    g = torch.Generator().manual_seed(5)
    batches = [(torch.stack([torch.randint(0, d, (512,), generator=g) for d in dims], 1).int(),
                torch.randint(0, 2, (512,), generator=g).float()) for _ in range(4)]
    crit = torch.nn.BCEWithLogitsLoss()
    m1, o1 = build()
    eager = []
    for x, y in batches:
        loss = crit(m1(x.to(DEV)), y.to(DEV))
        for o in o1:
            o.zero_grad()
        loss.backward()
        for o in o1:
            o.step()
        eager.append(float(loss))
    m2, o2 = build()
    init = {k: v.clone() for k, v in m2.state_dict().items()}
    step = GraphedTrainStep(m2, o2, crit, *batches[0])
    m2.load_state_dict(init)                       # the capture warm-up took real optimizer steps: rewind
    for o in o2:
        for st in o.state.values():
            for k, v in st.items():
                if torch.is_tensor(v):
                    v.zero_()
    graphed = [float(step(x.pin_memory(), y.pin_memory())) for x, y in batches]
    assert_close(np.asarray(graphed), np.asarray(eager), what="graph-replayed losses", atol_scale=1e-4)
    for (k, a), (_, b_) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if a.dtype == torch.float32:
            assert_close(b_.cpu().numpy(), a.cpu().numpy(), what=f"weights {k}", atol_scale=1e-4)


@pytest.mark.parametrize("emb_cfg", [{"name": "qr", "divider": 5}, {"name": "vanilla"}])
def test_side_stream_overlap_is_bit_identical_to_inline_execution(R, emb_cfg, monkeypatch):
    """The row sort started on a side stream after the forward gather, and the first layer's weight gradient run
    beside the embedding backward, must change nothing: same gradients bit for bit as the in-line order."""
    import recsys_benchmark_b200.functional as RF
    import recsys_benchmark_b200.linalg as LA

    def grads(early, side_dw):
        monkeypatch.setattr(RF, "EARLY_SORT", early)
        if not side_dw:
            monkeypatch.setattr(LA, "_side_dw_safe", lambda w: False)     # every weight gradient in line
        torch.manual_seed(9)
        m = R.get_ctr_model(CRITEO_DIMS, dict(num_factor=16, hidden_sizes=[400, 400], p_dropout=0.0,
                                              use_batchnorm=False, embedding_config=dict(emb_cfg))).to(DEV).train()
        g = torch.Generator().manual_seed(3)
        x = torch.stack([torch.randint(0, d, (4096,), generator=g) for d in CRITEO_DIMS], 1).int().to(DEV)
        y = torch.randint(0, 2, (4096,), generator=g).float().to(DEV)
        out = []
        for _ in range(2):                      # twice: the second pass reuses allocator blocks of the first
            for p in m.parameters():
                p.grad = None
            torch.nn.functional.binary_cross_entropy_with_logits(m(x), y).backward()
            torch.cuda.synchronize()
            out.append({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
        return out

    n0 = RF.SortedRows.constructed
    a = grads(True, True)
    n1 = RF.SortedRows.constructed
    b = grads(False, False)
    assert n1 - n0 == 2, "the early sort did not run (one SortedRows per training step expected)"
    assert RF.SortedRows.constructed == n1, "EARLY_SORT = False must sort inside the backward"
    for ga, gb in zip(a, b):
        assert ga.keys() == gb.keys()
        for k in ga:
            if k == "fc.weight":      # first-order gradient: float atomics across CTAs, order not fixed
                assert_close(ga[k].cpu().numpy(), gb[k].cpu().numpy(), what=k)
            else:
                assert torch.equal(ga[k], gb[k]), k


# ---------------- golden files at the production row width (D = 16, 39 fields: kernel templates <kind, 4, 4>) ---
def test_d16_vanilla_sparse_adam_and_fused_update(R):
    _run_steps(R, "d16_vanilla_sparse_adam", {"name": "vanilla", "sparse": True}, SPARSE_ADAM, 2)
    g, model, state = build_from_golden("d16_vanilla_sparse_adam", {"name": "vanilla", "sparse": True})
    model.load_state_dict(state, strict=True)
    model.to(DEV).train()
    opts = R.get_optimizers(model, dict(SPARSE_ADAM, fused_sparse=True))
    crit = torch.nn.BCEWithLogitsLoss()
    for s in range(2):
        loss = crit(model(_t(g[f"step{s}/x"])), _t(g[f"step{s}/y"]).float())
        for o in opts:
            o.zero_grad()
        loss.backward()
        for o in opts:
            o.step()
        assert_close(model.embedding._emb_module.weight.detach().cpu().numpy(),
                     g[f"step{s}/after/embedding._emb_module.weight"], what=f"d16 fused step {s}", atol_scale=5e-5)


def test_d16_qr_mult(R):
    _run_steps(R, "d16_qr_mult", {"name": "qr", "divider": 5}, ADAM, 1)


def test_d16_pep_feature_dim(R, tmp_path):
    _run_steps(R, "d16_pep_feature_dim", {"name": "pep", "checkpoint_weight_dir": str(tmp_path),
                                          "threshold_type": "feature_dim"}, ADAM, 1)


def test_d16_optembed(R):
    _run_steps(R, "d16_optembed", {"name": "deepfm_optembed"}, ADAM, 1, pre_step=_optembed_pre)


def test_gradient_accumulation_with_side_stream_dw_matches_inline(R, monkeypatch):
    """Two backward passes into the same .grad (gradient accumulation) and zero_grad(set_to_none=False): the first
    layer's weight gradient must not race its side-stream GEMM - results equal the in-line execution."""
    import recsys_benchmark_b200.linalg as LA

    def run(side):
        if not side:
            monkeypatch.setattr(LA, "_side_dw_safe", lambda w: False)
        torch.manual_seed(9)
        m = R.get_ctr_model(CRITEO_DIMS, dict(num_factor=16, hidden_sizes=[400, 400], p_dropout=0.0,
                                              use_batchnorm=False, embedding_config={"name": "qr", "divider": 5})
                            ).to(DEV).train()
        g = torch.Generator().manual_seed(3)
        xs = [torch.stack([torch.randint(0, d, (4096,), generator=g) for d in CRITEO_DIMS], 1).int().to(DEV)
              for _ in range(3)]
        y = torch.randint(0, 2, (4096,), generator=g).float().to(DEV)
        for p in m.parameters():
            p.grad = None
        for x in xs[:2]:                          # accumulate two micro-batches
            torch.nn.functional.binary_cross_entropy_with_logits(m(x), y).backward()
        acc = m._deep_branch[0].weight.grad.clone()
        m.zero_grad(set_to_none=False)            # .grad stays allocated: the next backward must add in place
        torch.nn.functional.binary_cross_entropy_with_logits(m(xs[2]), y).backward()
        torch.cuda.synchronize()
        return acc, m._deep_branch[0].weight.grad.clone()

    a = run(True)
    b = run(False)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
