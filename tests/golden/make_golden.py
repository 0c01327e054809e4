"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports chenxing1999/recsys-benchmark from /root/reference (with `lmdb`
stubbed: it is only touched when real dataset caches are opened), builds the
reference's own DeepFM / DCN_Mix with each embedding plugin on tiny field
sizes, runs forward / backward / optimizer steps on CPU in fp32, and stores
inputs, state dicts, outputs, gradients and post-step weights as
`tests/golden/<case>.npz`.  Those files are committed; the GPU box never sees
/root/reference.
"""
from __future__ import annotations

import copy
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("RSB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
sys.modules.setdefault("optuna", types.ModuleType("optuna"))

from loguru import logger  # noqa: E402

logger.remove()

from src.models import get_ctr_model  # noqa: E402
from src.models.deepfm import get_optimizers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

FIELD_DIMS = [7, 3, 11, 5, 2, 9]
D = 8
B = 24


def _np(t):
    return t.detach().cpu().numpy().copy()


def _batch(seed, field_dims=FIELD_DIMS, b=B, dtype=torch.int64):
    g = torch.Generator().manual_seed(seed)
    cols = [torch.randint(0, d, (b,), generator=g) for d in field_dims]
    x = torch.stack(cols, 1).to(dtype)
    y = torch.randint(0, 2, (b,), generator=g)
    return x, y


def _state(model, out, prefix="state/"):
    for k, v in model.state_dict().items():
        if torch.is_tensor(v):            # DHE's `_extra_state` is a dict ({"_prefix": int}); recorded separately
            out[prefix + k] = _np(v)


def _dense(g):
    if g is None:
        return None
    return _np(g.to_dense() if g.is_sparse else g)


def run_deepfm_case(name, emb_cfg, *, opt_cfg=None, tweak=None, steps=2, use_bn=True,
                    field_dims=FIELD_DIMS, d=D, int32_input=False, pre_forward=None, b=B, hidden=(16, 8)):
    """One golden file: eval logits, train-mode (dropout 0) logits, all grads,
    grad wrt embedding output / deep input, and `steps` optimizer steps."""
    torch.manual_seed(1234)
    cfg = dict(num_factor=d, hidden_sizes=list(hidden), p_dropout=0.0, use_batchnorm=use_bn,
               embedding_config=copy.deepcopy(emb_cfg))
    model = get_ctr_model(field_dims, cfg)
    if tweak is not None:
        tweak(model)
    out = {}
    out["field_dims"] = np.asarray(field_dims, dtype=np.int64)
    out["num_factor"] = np.asarray(d, dtype=np.int64)
    out["hidden_sizes"] = np.asarray(list(hidden), dtype=np.int64)
    _state(model, out)

    stash = {}

    def emb_hook(mod, inp, res):
        if res.requires_grad:
            res.retain_grad()
        stash["emb"] = res

    def deep_hook(mod, inp):
        if inp[0].requires_grad:
            inp[0].retain_grad()
        stash["deep_in"] = inp[0]

    h1 = model.embedding.register_forward_hook(emb_hook)
    h2 = model._deep_branch.register_forward_pre_hook(deep_hook)

    x, y = _batch(7, field_dims, b)
    if int32_input:
        x = x.int()
    out["x"], out["y"] = _np(x), _np(y)

    # eval-mode logits (BN running stats are the fresh 0/1)
    model.eval()
    if pre_forward is not None:
        pre_forward(model, "eval", out)
    with torch.no_grad():
        out["logits_eval"] = _np(model(x))
        out["emb_eval"] = _np(stash["emb"])

    opts = get_optimizers(model, opt_cfg) if opt_cfg is not None else []
    crit = torch.nn.BCEWithLogitsLoss()
    model.train()
    for s in range(steps):
        xs, ys = _batch(7 + s, field_dims, b)
        if int32_input:
            xs = xs.int()
        out[f"step{s}/x"], out[f"step{s}/y"] = _np(xs), _np(ys)
        if pre_forward is not None:
            pre_forward(model, f"step{s}", out)
        logits = model(xs)
        logits.retain_grad()
        loss = crit(logits, ys.float())
        for p in model.parameters():
            p.grad = None
        loss.backward()
        out[f"step{s}/logits"] = _np(logits)
        out[f"step{s}/loss"] = _np(loss)
        out[f"step{s}/g_logits"] = _np(logits.grad)
        out[f"step{s}/emb"] = _np(stash["emb"])
        out[f"step{s}/g_emb"] = _np(stash["emb"].grad)
        out[f"step{s}/g_deep"] = _np(stash["deep_in"].grad)
        for k, p in model.named_parameters():
            if p.grad is not None:
                out[f"step{s}/grad/{k}"] = _dense(p.grad)
        for o in opts:
            o.step()
        if opts:
            _state(model, out, prefix=f"step{s}/after/")
    h1.remove()
    h2.remove()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, len(out), "arrays")


def tweak_pep(scale_s=-2.0):
    def f(model):
        g = torch.Generator().manual_seed(99)
        emb = model.embedding
        with torch.no_grad():
            emb.emb.weight.uniform_(-0.5, 0.5, generator=g)
            emb.s.copy_(scale_s + 0.7 * torch.randn(emb.s.shape, generator=g).clamp_(-2, 2))
    return f


def tweak_cerp(model):
    g = torch.Generator().manual_seed(55)
    emb = model.embedding
    with torch.no_grad():
        emb.p_weight.uniform_(-0.5, 0.5, generator=g)
        emb.q_weight.uniform_(-0.5, 0.5, generator=g)
        emb.p_threshold.copy_(-2.0 + 0.7 * torch.randn(emb.p_threshold.shape, generator=g).clamp_(-2, 2))
        emb.q_threshold.copy_(-2.0 + 0.7 * torch.randn(emb.q_threshold.shape, generator=g).clamp_(-2, 2))


def tweak_optembed(model):
    g = torch.Generator().manual_seed(77)
    lo, span = (0.3, 0.8) if getattr(model.embedding._mask_e_module, "_norm", 1) == 1 else (0.2, 0.25)
    emb = model.embedding
    with torch.no_grad():
        emb._weight.uniform_(-0.2, 0.2, generator=g)
        if hasattr(emb._mask_e_module, "_t_param"):
            emb._mask_e_module._t_param.copy_(lo + span * torch.rand(emb._mask_e_module._t_param.shape, generator=g))


def optembed_pre_forward(model, tag, out):
    """Record the mask-D draw the reference is about to make
    (torch.randint(0, D, (B, F)) at deepfm_opt_embed.py:222-224)."""
    if tag == "eval":
        model.embedding.get_weight()  # fresh module: build the eval cache first
        out["eval_weight"] = _np(model.embedding._cur_weight)
        return
    seed = 4242 + int(tag[-1])
    torch.manual_seed(seed)
    k = torch.randint(0, model.embedding._hidden_size, size=(out[f"{tag}/x"].shape[0], model.embedding._num_field))
    out[f"{tag}/mask_d_idx"] = _np(k)
    torch.manual_seed(seed)


def run_dcn_case(name="dcn_mix"):
    torch.manual_seed(4321)
    field_dims = [6, 4, 9, 5]
    cfg = dict(name="dcn_mix", num_factor=4, hidden_sizes=[12], num_layers=2, num_experts=3, rank=5,
               p_dropout=0.0, compile_model=False, embedding_config={"name": "vanilla"})
    model = get_ctr_model(field_dims, cfg)
    with torch.no_grad():
        for b in model.cross_head.biases:
            b.normal_(0, 0.1)
    out = {"field_dims": np.asarray(field_dims, dtype=np.int64)}
    _state(model, out)
    stash = {}

    def pre(mod, inp):
        if inp[0].requires_grad:
            inp[0].retain_grad()
        stash["x0"] = inp[0]

    def post(mod, inp, res):
        if res.requires_grad:
            res.retain_grad()
        stash["xl"] = res

    model.cross_head.register_forward_pre_hook(pre)
    model.cross_head.register_forward_hook(post)
    x, y = _batch(11, field_dims, b=20)
    out["x"], out["y"] = _np(x), _np(y)
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = _np(model(x))
    model.train()
    logits = model(x)
    loss = torch.nn.BCEWithLogitsLoss()(logits, y.float())
    loss.backward()
    out["logits"] = _np(logits)
    out["cross_in"] = _np(stash["x0"])
    out["cross_out"] = _np(stash["xl"])
    out["g_cross_in"] = _np(stash["x0"].grad)
    out["g_cross_out"] = _np(stash["xl"].grad)
    for k, p in model.named_parameters():
        if p.grad is not None:
            out[f"grad/{k}"] = _dense(p.grad)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, len(out), "arrays")


def run_embedding_api_case():
    """1-D ids, get_weight, QR cat shapes, retrain masks: the plugin-level edges
    (tests/test_emb.py:124-164 in the reference check shapes only)."""
    from src.models.embeddings import get_embedding

    out = {}
    torch.manual_seed(5)
    fd = [5, 4, 6]
    ids1 = torch.tensor([0, 3, 14, 7, 7, 2])
    ids2 = torch.tensor([[0, 5, 9], [4, 8, 14], [2, 6, 11], [4, 5, 9]])
    out["ids1"], out["ids2"] = _np(ids1), _np(ids2)
    out["field_dims"] = np.asarray(fd)
    for op in ["mult", "add", "cat"]:
        emb = get_embedding({"name": "qr", "divider": 4, "operation": op}, fd, 6)
        for k, v in emb.state_dict().items():
            out[f"qr_{op}/state/{k}"] = _np(v)
        out[f"qr_{op}/out1"] = _np(emb(ids1))
        out[f"qr_{op}/out2"] = _np(emb(ids2))
        out[f"qr_{op}/get_weight"] = _np(emb.get_weight())
    emb = get_embedding({"name": "qr"}, fd, 6)  # default divider int(sqrt(15)) = 3
    out["qr_default/divider"] = np.asarray(emb._divider)
    out["qr_default/rows2"] = np.asarray(emb.emb2.weight.shape[0])

    # OptEmbed retrain mask
    emb = get_embedding({"name": "deepfm_optembed_retrain"}, fd, 6)
    torch.nn.init.uniform_(emb._weight, -1, 1)
    mask_e = torch.tensor([1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1])
    mask_d = torch.tensor([2, 5, 0])
    m = emb.init_mask(mask_e=mask_e, mask_d=mask_d)
    out["optretrain/weight"] = _np(emb._weight)
    out["optretrain/mask_e"], out["optretrain/mask_d"] = _np(mask_e), _np(mask_d)
    out["optretrain/mask"] = _np(m)
    out["optretrain/out2"] = _np(emb(ids2))
    sp, nnz = emb.get_sparsity(True)
    out["optretrain/sparsity"], out["optretrain/nnz"] = np.asarray(sp), np.asarray(nnz)

    # OptEmbed eval weight with explicit per-field mask_d + get_l_s + sparsity
    emb = get_embedding({"name": "deepfm_optembed"}, fd, 6)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        emb._weight.uniform_(-0.2, 0.2, generator=g)
        emb._mask_e_module._t_param.copy_(torch.tensor([0.5, 0.7, 0.55]))
    emb.eval()
    out["opteval/weight"] = _np(emb._weight)
    out["opteval/t"] = _np(emb._mask_e_module._t_param)
    out["opteval/w_plain"] = _np(emb.get_weight())
    out["opteval/w_maskd"] = _np(emb.get_weight(torch.tensor([1, 4, 2])))
    out["opteval/l_s"] = _np(emb.get_l_s())
    sp, nnz = emb.get_sparsity(True)
    out["opteval/sparsity"], out["opteval/nnz"] = np.asarray(sp), np.asarray(nnz)
    out["opteval/mask_e"] = _np(emb.get_mask_e())
    out["opteval/submask"] = _np(emb.get_submask())

    # PEP sparsity bookkeeping
    with tempfile.TemporaryDirectory() as td:
        emb = get_embedding({"name": "pep", "checkpoint_weight_dir": td, "threshold_type": "feature_dim"},
                            fd, 6, field_name="deepfm")
        with torch.no_grad():
            emb.emb.weight.uniform_(-0.5, 0.5, generator=g)
            emb.s.copy_(-1.5 + torch.randn(emb.s.shape, generator=g))
        out["pep/weight"], out["pep/s"] = _np(emb.emb.weight), _np(emb.s)
        sp, nnz = emb.get_sparsity(True)
        out["pep/sparsity"], out["pep/nnz"] = np.asarray(sp), np.asarray(nnz)
        out["pep/get_weight"] = _np(emb.get_weight())
        out["pep/out1"] = _np(emb(ids1))
    np.savez_compressed(os.path.join(HERE, "embedding_api.npz"), **out)
    print("wrote embedding_api", len(out), "arrays")


def run_pruned_csr_case():
    """Inference path: the reference's PrunedEmbedding (CSR + numba CPU kernel,
    src/models/embeddings/pruned_embedding.py) standing in for the embedding of an eval-mode DeepFM, the way
    scripts/deepfm/infer_deepfm.py:138-153 builds it."""
    from src.models.embeddings.pruned_embedding import PrunedEmbedding

    torch.manual_seed(21)
    model = get_ctr_model(list(FIELD_DIMS), dict(num_factor=D, hidden_sizes=[16, 8], p_dropout=0.1, use_batchnorm=True,
                                                 embedding_config={"name": "vanilla"}))
    g = torch.Generator().manual_seed(21)
    w = model.embedding.get_weight().detach().clone()
    keep = torch.rand(w.shape, generator=g) > 0.8          # ~80 % pruned
    keep[3] = False                                         # an all-zero row
    keep[5] = True                                          # a fully dense row
    keep[len(w) - 1] = False
    keep[len(w) - 1, D - 1] = True                          # last row: only the last dim
    w = w * keep
    with torch.no_grad():
        model.embedding._emb_module.weight.copy_(w)
    model.eval()
    out = {"field_dims": np.asarray(FIELD_DIMS, dtype=np.int64)}
    _state(model, out)
    x, _ = _batch(500)
    x32, _ = _batch(501, dtype=torch.int32)
    offsets = torch.tensor([0] + FIELD_DIMS[:-1]).cumsum(0)
    pruned = PrunedEmbedding.from_other_emb(model.embedding)
    out["csr/values"] = np.asarray(pruned.values).copy()
    out["csr/crow_indices"] = np.asarray(pruned.crow_indices).copy()
    out["csr/col_indices"] = np.asarray(pruned.col_indices).copy()
    out["x"] = _np(x)
    out["x_int32"] = _np(x32)
    with torch.no_grad():
        out["emb"] = _np(pruned(x + offsets))
        out["emb_1d"] = _np(pruned((x + offsets)[:, 2].contiguous()))
        out["weight_dense"] = _np(pruned.get_weight())
        model.embedding = pruned
        out["logits"] = _np(model(x))
        out["logits_int32"] = _np(model(x32))
    np.savez_compressed(os.path.join(HERE, "pruned_csr.npz"), **out)
    print("wrote pruned_csr", len(out), "arrays")


def run_dhe_cases():
    """Deep hash embedding (src/models/embeddings/dh_embedding.py): universal-hash codes,
    the cached table, and DeepFM train steps through the hash -> MLP encoder."""
    from src.models.embeddings.dh_embedding import DHEmbedding

    adam = dict(learning_rate=1e-2, weight_decay=1e-4)
    for name, cfg, prefix in [
        ("deepfm_dhe", {"name": "dhe", "inp_size": 32, "hidden_sizes": [16]}, 0),
        ("deepfm_dhe_v2", {"name": "dhe", "inp_size": 24, "hidden_sizes": [], "use_bn": 1, "compute_v2": True}, 1000),
        ("deepfm_dhe_nobn", {"name": "dhe", "inp_size": 32, "hidden_sizes": [12, 16], "use_bn": 0}, 37),
    ]:
        DHEmbedding.COUNTER = prefix

        def record(model, tag, out, _prefix=prefix):
            if tag != "eval":
                return
            e = model.embedding
            assert e._prefix == _prefix
            out["dhe/prefix"] = np.asarray(e._prefix, dtype=np.int64)
            out["dhe/cache"] = _np(e._cache)
            ids = torch.tensor([0, 1, 5, len(e._cache) - 1, 123456789, 2 ** 31 + 7])
            out["dhe/ids"] = _np(ids)
            out["dhe/hash_batch"] = _np(e._get_universal_hash_batch(ids))

        run_deepfm_case(name, cfg, opt_cfg=adam, steps=2, pre_forward=record)


def run_cerp_retrain_case():
    """CERP retrain (cerp_embedding.py:209-378): P, Q re-initialised from `initial.pth`, bool masks from the
    searched checkpoint `target.pth` (|w| - sigmoid(threshold) > 0), masked tables gathered and added."""
    bucket = 5
    g = torch.Generator().manual_seed(31)
    tgt = {"q_weight": torch.empty(bucket, D).uniform_(-0.6, 0.6, generator=g),
           "p_weight": torch.empty(bucket, D).uniform_(-0.6, 0.6, generator=g),
           "q_threshold": -1.2 + torch.randn(bucket, D, generator=g),
           "p_threshold": -1.2 + torch.randn(bucket, D, generator=g)}
    init = {"q_weight": torch.empty(bucket, D).uniform_(-0.5, 0.5, generator=g),
            "p_weight": torch.empty(bucket, D).uniform_(-0.5, 0.5, generator=g)}
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "deepfm"))
        torch.save(tgt, os.path.join(td, "deepfm", "target.pth"))
        torch.save(init, os.path.join(td, "deepfm", "initial.pth"))
        run_deepfm_case("deepfm_cerp_retrain", {"name": "cerp_retrain", "checkpoint_weight_dir": td,
                                                "bucket_size": bucket},
                        opt_cfg=dict(learning_rate=1e-2, weight_decay=1e-4), steps=2)
    np.savez_compressed(os.path.join(HERE, "cerp_retrain_ckpt.npz"),
                        **{"target/" + k: _np(v) for k, v in tgt.items()},
                        **{"initial/" + k: _np(v) for k, v in init.items()})


# A second set at the PRODUCTION row width (D = 16, the kernels' <kind, V=4, LPR=4> instantiation) and the Criteo
# field count (39 small fields): the D = 8 files above exercise <kind, 4, 2> only.
D16_DIMS = [7, 3, 11, 5, 2, 9, 4, 13, 6, 8, 3, 5, 7, 19, 12, 31, 23, 6, 4, 17, 9, 2, 29, 14, 37, 10, 5, 21, 33, 3, 11,
            8, 2, 27, 4, 6, 25, 9, 15]


def run_d16_cases():
    kw = dict(field_dims=D16_DIMS, d=16, b=64, hidden=(32, 16))
    adam = dict(learning_rate=1e-2, weight_decay=1e-4)
    run_deepfm_case("d16_vanilla_sparse_adam", {"name": "vanilla", "sparse": True},
                    opt_cfg=dict(learning_rate=1e-2, weight_decay=1e-4, sparse=True), steps=2, **kw)
    run_deepfm_case("d16_qr_mult", {"name": "qr", "divider": 5}, opt_cfg=adam, use_bn=False, steps=1, **kw)
    with tempfile.TemporaryDirectory() as td:
        run_deepfm_case("d16_pep_feature_dim", {"name": "pep", "checkpoint_weight_dir": td,
                                                "threshold_type": "feature_dim"},
                        opt_cfg=adam, tweak=tweak_pep(), steps=1, **kw)
    run_deepfm_case("d16_optembed", {"name": "deepfm_optembed"}, opt_cfg=adam, tweak=tweak_optembed, steps=1,
                    pre_forward=optembed_pre_forward, **kw)


def main():
    if len(sys.argv) > 1:          # regenerate only the named extra cases
        for name in sys.argv[1:]:
            {"pruned_csr": run_pruned_csr_case, "dhe": run_dhe_cases, "cerp_retrain": run_cerp_retrain_case,
             "d16": run_d16_cases}[name]()
        return
    run_d16_cases()
    run_pruned_csr_case()
    run_dhe_cases()
    run_cerp_retrain_case()
    adam = dict(learning_rate=1e-2, weight_decay=1e-4)
    sparse_adam = dict(learning_rate=1e-2, weight_decay=1e-4, sparse=True)
    sparse_sgd = dict(learning_rate=1e-1, weight_decay=1e-4, sparse=True, optimizer="sgd")

    run_deepfm_case("deepfm_vanilla_adam", {"name": "vanilla"}, opt_cfg=adam)
    run_deepfm_case("deepfm_vanilla_sparse_adam", {"name": "vanilla", "sparse": True}, opt_cfg=sparse_adam, steps=3)
    run_deepfm_case("deepfm_vanilla_sparse_sgd", {"name": "vanilla", "sparse": True}, opt_cfg=sparse_sgd)
    run_deepfm_case("deepfm_vanilla_int32", {"name": "vanilla"}, opt_cfg=None, steps=1, int32_input=True)
    for op in ["mult", "add", "cat"]:
        run_deepfm_case(f"deepfm_qr_{op}", {"name": "qr", "divider": 4, "operation": op}, opt_cfg=adam, use_bn=False)
    with tempfile.TemporaryDirectory() as td:
        for tt in ["feature_dim", "feature", "dimension", "global"]:
            run_deepfm_case(f"deepfm_pep_{tt}",
                            {"name": "pep", "checkpoint_weight_dir": td, "threshold_type": tt},
                            opt_cfg=adam, tweak=tweak_pep(), steps=1)
        # PEP retrain: needs a finished PEP checkpoint {dir}/deepfm/{sparsity}.pth
        torch.manual_seed(8)
        n = sum(FIELD_DIMS)
        g = torch.Generator().manual_seed(8)
        fin_w = torch.empty(n, D).uniform_(-0.5, 0.5, generator=g)
        fin_s = -1.0 + torch.randn(n, D, generator=g)
        os.makedirs(os.path.join(td, "deepfm"), exist_ok=True)
        torch.save({"emb.weight": fin_w, "s": fin_s}, os.path.join(td, "deepfm", "0.5.pth"))
        for sp in [False, True]:
            run_deepfm_case("deepfm_pep_retrain" + ("_sparse" if sp else ""),
                            {"name": "pep_retrain", "checkpoint_weight_dir": td, "sparsity": 0.5, "sparse": sp},
                            opt_cfg=(sparse_adam if sp else adam), steps=1)
        np.savez_compressed(os.path.join(HERE, "pep_retrain_ckpt.npz"), weight=_np(fin_w), s=_np(fin_s))
    run_deepfm_case("deepfm_optembed", {"name": "deepfm_optembed"}, opt_cfg=adam, tweak=tweak_optembed,
                    steps=2, pre_forward=optembed_pre_forward)
    run_deepfm_case("deepfm_optembed_l2", {"name": "deepfm_optembed", "norm": 2}, opt_cfg=None,
                    tweak=tweak_optembed, steps=1, pre_forward=optembed_pre_forward)
    run_deepfm_case("deepfm_optembed_d", {"name": "deepfm_optembed_d"}, opt_cfg=None, tweak=tweak_optembed,
                    steps=1, pre_forward=optembed_pre_forward)
    run_deepfm_case("deepfm_cerp", {"name": "cerp", "bucket_size": 5, "threshold_init": -2.0,
                                    "threshold_init_method": "uniform"}, opt_cfg=adam, tweak=tweak_cerp, steps=2)
    run_dcn_case()
    run_embedding_api_case()


if __name__ == "__main__":
    main()
