"""CPU tests of the host-side mirror of the reference API: registry / factory semantics,
state-dict compatibility with the reference (from the golden files), optimizer grouping,
and that the product path refuses to run without CUDA (no CPU fallback)."""
import numpy as np
import pytest
import torch

import recsys_benchmark_b200 as R
from tests.helpers import load_golden, sub

CASES = {
    "deepfm_vanilla_adam": {"name": "vanilla"},
    "deepfm_vanilla_sparse_adam": {"name": "vanilla", "sparse": True},
    "deepfm_qr_mult": {"name": "qr", "divider": 4, "operation": "mult"},
    "deepfm_qr_add": {"name": "qr", "divider": 4, "operation": "add"},
    "deepfm_qr_cat": {"name": "qr", "divider": 4, "operation": "cat"},
    "deepfm_optembed": {"name": "deepfm_optembed"},
    "deepfm_optembed_l2": {"name": "deepfm_optembed", "norm": 2},
    "deepfm_optembed_d": {"name": "deepfm_optembed_d"},
    "deepfm_cerp": {"name": "cerp", "bucket_size": 5, "threshold_init": -2.0, "threshold_init_method": "uniform"},
}


def build_from_golden(name, emb_cfg, tmp_path=None, **model_kw):
    g = load_golden(name)
    fd = [int(v) for v in g["field_dims"]]
    st = sub(g, "state/")
    use_bn = any(k.endswith("running_mean") for k in st)
    d = int(g["num_factor"]) if "num_factor" in g else 8
    hidden = [int(v) for v in g["hidden_sizes"]] if "hidden_sizes" in g else [16, 8]
    cfg = dict(num_factor=d, hidden_sizes=hidden, p_dropout=0.0, use_batchnorm=use_bn,
               embedding_config=dict(emb_cfg))
    cfg.update(model_kw)
    model = R.get_ctr_model(fd, cfg)
    return g, model, {k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}


@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dict_matches_reference(name):
    g, model, state = build_from_golden(name, CASES[name])
    ours = model.state_dict()
    assert sorted(ours.keys()) == sorted(state.keys())
    for k, v in state.items():
        assert tuple(ours[k].shape) == tuple(v.shape), k
        assert ours[k].dtype == v.dtype, k
    model.load_state_dict(state, strict=True)
    ref_grad_keys = {k[len("step0/grad/"):] for k in g if k.startswith("step0/grad/")}
    trainable = {k for k, p in model.named_parameters() if p.requires_grad}
    assert ref_grad_keys <= trainable


@pytest.mark.parametrize("tt", ["feature_dim", "feature", "dimension", "global"])
def test_pep_state_dict(tt, tmp_path):
    g, model, state = build_from_golden(
        f"deepfm_pep_{tt}", {"name": "pep", "checkpoint_weight_dir": str(tmp_path), "threshold_type": tt})
    model.load_state_dict(state, strict=True)
    assert (tmp_path / "deepfm").is_dir()  # constructor side effect kept (pep_embedding.py:73-75)
    assert model.embedding.checkpoint_weight_dir.endswith("deepfm")
    assert model.embedding.sparsity == [0.8, 0.9, 0.99] and model.embedding._cur_min_spar_idx == 0


def test_pep_retrain_state_dict_and_mask(tmp_path):
    ck = load_golden("pep_retrain_ckpt")
    (tmp_path / "deepfm").mkdir()
    torch.save({"emb.weight": torch.from_numpy(ck["weight"]), "s": torch.from_numpy(ck["s"])},
               tmp_path / "deepfm" / "0.5.pth")
    g, model, state = build_from_golden(
        "deepfm_pep_retrain", {"name": "pep_retrain", "checkpoint_weight_dir": str(tmp_path), "sparsity": 0.5})
    assert model.embedding.mask.dtype == torch.bool and not model.embedding.mask.requires_grad
    np.testing.assert_array_equal(model.embedding.mask.numpy(), g["state/embedding.mask"])
    model.load_state_dict(state, strict=True)


def test_factory_semantics():
    with pytest.raises(NotImplementedError):
        R.get_embedding({"name": "nope"}, [3, 4], 8)
    with pytest.raises(AssertionError):
        R.get_embedding({"name": "vanilla"}, [3, 4], 8, mode="bad")
    cfg = {"name": "deepfm_optembed_d"}
    emb = R.get_embedding(cfg, [3, 4], 8)
    assert cfg == {"name": "deepfm_optembed_d"}  # config not mutated
    assert isinstance(emb._mask_e_module, torch.nn.Identity)
    assert [k for k, _ in emb.named_parameters()] == ["_weight"]
    qr = R.get_embedding({"name": "qr"}, [5, 4, 6], 6)
    assert qr._divider == 3 and qr.emb2.weight.shape == (5, 6)
    qr = R.get_embedding({"name": "qr", "divider": 4, "operation": "cat"}, [5, 4, 6], 6)
    assert qr.emb1.weight.shape == (4, 3) and qr.emb2.weight.shape == (4, 3)
    v = R.get_embedding({"name": "vanilla", "sparse": True}, 10, 4)
    assert v._emb_module.sparse and v.get_weight().shape == (10, 4)


def test_get_optimizers_grouping():
    model = R.get_ctr_model([3, 4, 5], dict(num_factor=4, hidden_sizes=[8],
                                            embedding_config={"name": "vanilla", "sparse": True}))
    opts = R.get_optimizers(model, dict(learning_rate=1e-3, weight_decay=1e-6))
    assert len(opts) == 1 and isinstance(opts[0], torch.optim.Adam)
    opts = R.get_optimizers(model, dict(learning_rate=1e-3, weight_decay=1e-6, sparse=True))
    assert isinstance(opts[0], torch.optim.SparseAdam) and isinstance(opts[1], torch.optim.Adam)
    emb_ids = {id(p) for p in model.embedding.parameters()}
    assert {id(p) for g in opts[0].param_groups for p in g["params"]} == emb_ids
    rest = {id(p) for g in opts[1].param_groups for p in g["params"]}
    assert id(model.fc.weight) in rest and not (rest & emb_ids)   # fc stays under dense Adam
    opts = R.get_optimizers(model, dict(learning_rate=1e-3, weight_decay=1e-6, sparse=True, optimizer="sgd"))
    assert len(opts) == 1 and opts[0].param_groups[0]["weight_decay"] == 0
    opts = R.get_optimizers(model, dict(learning_rate=1e-3, weight_decay=0, sparse=True, fused_sparse=True))
    assert isinstance(opts[0], R.FusedSparseAdam) and model.embedding._rsb_fused_opt is opts[0]
    with pytest.raises(ValueError):
        R.get_optimizers(model, dict(learning_rate=1e-3, weight_decay=0, optimizer="lamb"))


def test_cpu_tensors_are_refused_no_fallback():
    model = R.get_ctr_model([3, 4, 5], dict(num_factor=4, hidden_sizes=[8]))
    x = torch.zeros(2, 3, dtype=torch.long)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        model(x)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        model.embedding(x)


def test_dcn_mix_state_dict_and_orig_mod_prefix():
    g = load_golden("dcn_mix")
    fd = [int(v) for v in g["field_dims"]]
    cfg = dict(name="dcn_mix", num_factor=4, hidden_sizes=[12], num_layers=2, num_experts=3, rank=5, p_dropout=0.0,
               compile_model=True, embedding_config={"name": "vanilla"})
    model = R.get_ctr_model(fd, cfg)
    assert cfg["compile_model"] is True and "name" not in cfg
    state = {k: torch.from_numpy(np.asarray(v)) for k, v in sub(g, "state/").items()}
    assert sorted(model.state_dict()) == sorted(state)
    model.load_state_dict(state, strict=True)
    ck = {"field_dims": fd, "state_dict": {"_orig_mod." + k: v for k, v in state.items()},
          "model_config": dict(num_factor=4, hidden_sizes=[12], num_layers=2, num_experts=3, rank=5, p_dropout=0.0,
                               compile_model=True, embedding_config={"name": "vanilla"})}
    m2 = R.DCN_Mix.load(ck)
    torch.testing.assert_close(m2.cross_head.gates, model.cross_head.gates)


def test_dcn_head_torch_math_matches_reference_golden_on_cpu():
    """The restructured cross layer (single GEMM over U as [E*r, Dm]) equals the reference."""
    g = load_golden("dcn_mix")
    st = sub(g, "state/")
    head = R.DCN_MixHead(3, 2, 5, 16)
    head.load_state_dict({k[len("cross_head."):]: torch.from_numpy(v) for k, v in st.items()
                          if k.startswith("cross_head.")})
    x0 = torch.from_numpy(g["cross_in"]).requires_grad_(True)
    out = head(x0)
    torch.testing.assert_close(out, torch.from_numpy(g["cross_out"]), rtol=1e-5, atol=1e-6)
    out.backward(torch.from_numpy(g["g_cross_out"]))
    torch.testing.assert_close(x0.grad, torch.from_numpy(g["g_cross_in"]), rtol=1e-4, atol=1e-7)
    for n in ["V", "C", "U", "biases"]:
        for l in range(2):
            torch.testing.assert_close(getattr(head, n)[l].grad, torch.from_numpy(g[f"grad/cross_head.{n}.{l}"]),
                                       rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(head.gates.grad, torch.from_numpy(g["grad/cross_head.gates"]), rtol=1e-4, atol=1e-7)


# ------------------------------------------------------- pruned CSR (f-4) ---
def test_pruned_embedding_host_side_matches_reference_layout():
    g = load_golden("pruned_csr")
    w = torch.from_numpy(np.asarray(g["state/embedding._emb_module.weight"]))
    emb = R.PrunedEmbedding.from_weight(w)
    np.testing.assert_array_equal(emb.values.numpy(), g["csr/values"])
    np.testing.assert_array_equal(emb.crow_indices.numpy(), g["csr/crow_indices"])
    np.testing.assert_array_equal(emb.col_indices.numpy(), g["csr/col_indices"])
    np.testing.assert_array_equal(emb.get_weight().numpy(), g["weight_dense"])
    assert emb.state_dict() == {} and not emb.is_cuda                 # nothing persistent, like the reference
    assert emb.get_num_params() == int((w != 0).sum())
    small = R.PrunedEmbedding.from_weight(w, compact=True)
    assert small.crow_indices.dtype == torch.int32 and small.col_indices.dtype == torch.uint8
    np.testing.assert_array_equal(small.get_weight().numpy(), g["weight_dense"])
    van = R.get_embedding({"name": "vanilla"}, [int(v) for v in g["field_dims"]], w.shape[1])
    with torch.no_grad():
        van._emb_module.weight.copy_(w)
    np.testing.assert_array_equal(R.PrunedEmbedding.from_other_emb(van).values.numpy(), g["csr/values"])
    with pytest.raises(RuntimeError):                                  # no CPU fallback for the gather
        emb(torch.zeros(2, 3, dtype=torch.int64))


def test_pruned_embedding_canonicalises_unsorted_csr():
    crow = torch.tensor([0, 2, 2, 3])
    col = torch.tensor([3, 1, 0])                                      # row 0 has its columns out of order
    val = torch.tensor([1.0, 2.0, 3.0])
    csr = torch.sparse_csr_tensor(crow, col, val, size=(3, 4))
    emb = R.PrunedEmbedding.from_weight(csr)
    assert emb.col_indices.tolist() == [1, 3, 0] and emb.values.tolist() == [2.0, 1.0, 3.0]
    np.testing.assert_array_equal(emb.get_weight().numpy(), csr.to_dense().numpy())


# ------------------------------------------------------- deep hash embedding (f-3) ---
def test_dhe_state_dict_and_counter_semantics():
    g = load_golden("deepfm_dhe")
    fd = [int(v) for v in g["field_dims"]]
    R.DHEmbedding.COUNTER = 0
    model = R.get_ctr_model(fd, dict(num_factor=8, hidden_sizes=[16, 8], p_dropout=0.0, use_batchnorm=True,
                                     embedding_config={"name": "dhe", "inp_size": 32, "hidden_sizes": [16]}))
    ours = model.state_dict()
    ref = sub(g, "state/")
    assert sorted(k for k in ours if k != "embedding._extra_state") == sorted(ref.keys())
    for k, v in ref.items():
        assert tuple(ours[k].shape) == tuple(v.shape) and ours[k].numpy().dtype == v.dtype, k
    # same seeded draws as the reference constructor -> identical hash coefficients before any state dict is loaded
    for k in ("_slopes", "_bias", "_primes_choices"):
        np.testing.assert_array_equal(ours["embedding." + k].numpy(), ref["embedding." + k])
    assert ours["embedding._extra_state"] == {"_prefix": 0} and R.DHEmbedding.COUNTER == sum(fd)
    second = R.DHEmbedding(fd, 8, None, 32, [16])                 # class-level counter keeps tables apart
    assert second._prefix == sum(fd)
    hs = [16]
    R.DHEmbedding(fd, 8, None, 32, hs)
    with pytest.raises(NotImplementedError):
        R.DHEmbedding(fd, 8, use_universal_hash=False)
    with pytest.raises(RuntimeError):
        model.embedding.encode(torch.zeros(3, dtype=torch.int64))  # CPU tensors: no fallback


# ------------------------------------------------------------ CERP retrain (f-1) ---
def make_cerp_retrain_dir(tmp_path):
    ck = load_golden("cerp_retrain_ckpt")
    (tmp_path / "deepfm").mkdir(exist_ok=True)
    for part in ("target", "initial"):
        torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in sub(ck, part + "/").items()},
                   tmp_path / "deepfm" / f"{part}.pth")
    return {"name": "cerp_retrain", "checkpoint_weight_dir": str(tmp_path), "bucket_size": 5}


def test_cerp_retrain_state_dict_and_masks(tmp_path):
    g, model, state = build_from_golden("deepfm_cerp_retrain", make_cerp_retrain_dir(tmp_path))
    ours = model.state_dict()
    assert sorted(ours.keys()) == sorted(state.keys())
    for k in ("embedding.q_mask", "embedding.p_mask", "embedding.q_weight", "embedding.p_weight"):
        assert ours[k].dtype == state[k].dtype and torch.equal(ours[k], state[k]), k   # loaded from the checkpoint dir
    assert not model.embedding.q_mask.requires_grad and model.embedding.q_weight.requires_grad
    assert model.embedding.get_num_params() == int(state["embedding.q_mask"].sum() + state["embedding.p_mask"].sum())
    with pytest.raises(AssertionError):
        R.get_embedding({"name": "cerp_retrain", "checkpoint_weight_dir": str(tmp_path / "nope"), "bucket_size": 5},
                        [3, 4], 8, field_name="deepfm")


# ------------------------------------------------------------ checkpoints ---
def test_checkpoint_round_trip_like_the_reference_scripts(tmp_path):
    """scripts/deepfm/train_deepfm.py:204-210 writes {"state_dict", "model_config", "field_dims"}; DeepFM.load /
    DCN_Mix.load / load_ctr_model read it back (src/models/deepfm.py:126-133, src/models/__init__.py:92-131);
    save_model_checkpoint writes the embedding's own state dict to {dir}/{field}/{name}.pth."""
    fd = [7, 3, 11]
    cfg = dict(num_factor=8, hidden_sizes=[16, 8], p_dropout=0.1, use_batchnorm=True,
               embedding_config={"name": "qr", "divider": 4})
    torch.manual_seed(0)
    model = R.get_ctr_model(fd, dict(cfg))
    ck = {"state_dict": model.state_dict(), "model_config": dict(cfg), "field_dims": fd}
    path = str(tmp_path / "deepfm_checkpoint.pth")
    torch.save(ck, path)
    for loaded in (R.DeepFM.load(path), R.load_ctr_model({"name": "deepfm"}, torch.load(path))):
        assert isinstance(loaded, R.DeepFM)
        for k, v in model.state_dict().items():
            assert torch.equal(loaded.state_dict()[k], v), k
    bare = R.DeepFM.load(ck, strict=False, empty_embedding=True)      # infer_deepfm.py:150 builds the table itself
    assert not hasattr(bare, "embedding")
    R.save_ctr_checkpoint(model, str(tmp_path), "target")
    saved = torch.load(tmp_path / "deepfm" / "target.pth")
    assert sorted(saved.keys()) == sorted(model.embedding.state_dict().keys())

    dcfg = dict(num_factor=4, hidden_sizes=[12], num_layers=2, num_experts=3, rank=5, p_dropout=0.0,
                embedding_config={"name": "vanilla"})
    dcn = R.get_ctr_model(fd, dict(dcfg, name="dcn_mix", compile_model=True))
    assert isinstance(dcn, R.DCN_Mix)
    compiled_keys = {"_orig_mod." + k: v for k, v in dcn.state_dict().items()}   # what a torch.compile'd reference saves
    back = R.DCN_Mix.load({"state_dict": compiled_keys, "model_config": dict(dcfg, compile_model=True),
                           "field_dims": fd})
    for k, v in dcn.state_dict().items():
        assert torch.equal(back.state_dict()[k], v), k
    R.save_ctr_checkpoint(dcn, str(tmp_path), "init")
    assert (tmp_path / "dcn" / "init.pth").exists()
    with pytest.raises(NotImplementedError):
        R.save_ctr_checkpoint(torch.nn.Linear(2, 2), str(tmp_path))


# ------------------------------------------------------------ GEMM host logic ---
def test_gemm_shape_gate_and_split_k_choice():
    import recsys_benchmark_b200.linalg as LA
    from recsys_benchmark_b200 import _lib

    # fp32 results are stored 128 bits at a time: N and the output leading dimension multiples of 4; operands are
    # re-laid out as bf16 planes, so K and M are free; A^T B^T is not provided
    assert LA.gemm_supported(65536, 400, 624, 624, 624, 400, False, True)
    assert LA.gemm_supported(65536, 400, 622, 622, 622, 400, False, True)
    assert not LA.gemm_supported(64, 5, 64, 64, 5, 5, False, False)          # the toy DCN rank (r = 5): library GEMM
    assert not LA.gemm_supported(64, 64, 64, 64, 64, 64, True, True)
    # weight-gradient GEMMs (K = batch): the library splits the reduction so that (tile, split) units fill the 148
    # SMs in one wave, every split at least 256 deep; the workspace query reports the partial buffers it needs
    lib = _lib.load()
    for m, n in ((400, 624), (400, 400), (256, 352)):
        ws = lib.rsb_gemm_planes_workspace_bytes(m, n, 65536, 1, 0)
        splits = (ws - 256) // (m * n * 4)
        tiles = -(-m // 128) * -(-n // 256)
        assert 148 // 2 < tiles * splits <= 148 and 65536 // splits >= 256, (m, n, splits)
    assert lib.rsb_gemm_planes_workspace_bytes(65536, 400, 624, 1, 0) == 256   # enough tiles: no split
    assert lib.rsb_gemm_planes_workspace_bytes(400, 400, 512, 1, 0) <= 256 + 2 * 400 * 400 * 4   # short reductions
    x = torch.randn(8, 6)
    lin = torch.nn.Linear(6, 3)
    torch.testing.assert_close(LA.linear(x, lin.weight, lin.bias), lin(x))    # off-path CPU tensors: plain F.linear
