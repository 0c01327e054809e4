"""the reference's trainers run UNCHANGED on the B200-native classes.

`recsys_benchmark_b200.install_into_reference()` rebinds the reference's registry; then the reference's own factory
(`src.models.get_ctr_model`), its own `get_optimizers`, and its own loops - `src/trainer/deepfm.py::train_epoch` /
`validate_epoch`, `scripts/deepfm/train_deepfm_pep.py::train_epoch` (clip_grad=100, get_sparsity + train_callback
every log step) and `scripts/deepfm/train_deepfm_optembed.py::train_epoch` (alpha * get_l_s(), Adam + the SGD
`t_param` group) - are executed verbatim for 20 steps on synthetic loaders on the GPU.  The same loops are then run
on the reference's own classes from the same state dict: the epoch losses must agree.
The reference is the copy staged under baseline/_ref (tests/refimport.py), not /root/reference."""
import copy

import pytest
import torch

from tests import refimport

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refimport.available(), reason="baseline/_ref not staged")]
DEV = "cuda:0"
DIMS = [49, 101, 126, 45, 223, 118, 84, 76, 95, 9, 30, 40, 75, 1458, 555, 19394, 13880, 306, 19, 11970, 634,
        4, 42646, 5178, 19277, 3175, 27, 11422, 18107, 11, 4654, 2032, 5, 18965, 18, 16, 59697, 86, 45571]
KDD = [60000] * 8 + [40000] * 3
STEPS, BATCH = 20, 1024


def _loader(dims, steps=STEPS, batch=BATCH, seed=0):
    """What the reference's DataLoader yields (criteo_torchfm.py:88-93): int32 ids [B,F] without offsets, labels [B].
    Labels depend on the ids so that 20 steps of training visibly reduce the loss."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        x = torch.stack([torch.randint(0, d, (batch,), generator=g) for d in dims], 1).int()
        y = ((x[:, 0] + x[:, 3] + x[:, 5]) % 2).long()
        out.append((x, y))
    return out


def _cfg(emb, p_dropout=0.0, **kw):
    c = dict(num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=p_dropout, use_batchnorm=True,
             embedding_config=dict(emb))
    c.update(kw)
    return c


CASES = {
    "vanilla_adam": (DIMS, {"name": "vanilla"}, dict(learning_rate=1e-3, weight_decay=1e-6)),
    "vanilla_sparse_adam": (DIMS, {"name": "vanilla", "sparse": True},
                            dict(learning_rate=1e-3, weight_decay=1e-6, sparse=True)),
    "qr5": (DIMS, {"name": "qr", "divider": 5}, dict(learning_rate=1e-3, weight_decay=1e-6)),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_unmodified_train_and_validate_epoch_on_installed_classes(case):
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200 as R

    dims, emb, opt_cfg = CASES[case]
    refimport.activate()
    import src.models as ref_models
    import src.models.deepfm as ref_deepfm
    import src.trainer.deepfm as ref_trainer

    torch.manual_seed(2023)
    ref_model = ref_models.get_ctr_model(dims, _cfg(emb))
    assert type(ref_model).__module__.startswith("src.")
    state = copy.deepcopy(ref_model.state_dict())
    loader = _loader(dims)
    ref_out = ref_trainer.train_epoch(loader, ref_model, ref_deepfm.get_optimizers(ref_model, dict(opt_cfg)), DEV,
                                      log_step=10)
    ref_val = ref_trainer.validate_epoch(loader[:4], ref_model, DEV)

    with refimport.installed() as mods:
        model = mods["src.models"].get_ctr_model(dims, _cfg(emb))          # the reference's factory, our classes
        assert isinstance(model, R.DeepFM) and isinstance(model.embedding, R.IEmbedding)
        model.load_state_dict(state, strict=True)
        opts = mods["src.models.deepfm"].get_optimizers(model, dict(opt_cfg))
        out = mods["src.trainer.deepfm"].train_epoch(loader, model, opts, DEV, log_step=10)   # unchanged loop
        val = mods["src.trainer.deepfm"].validate_epoch(loader[:4], model, DEV)
    assert ref_models.DeepFM.__module__.startswith("src."), "registry not restored"
    # 20 Adam steps: fp32 rounding differences between two correct implementations (GEMM, BatchNorm statistics) are
    # amplified by g / (sqrt(v) + eps); the epoch-mean losses agree to a few 1e-4 relative
    assert out["loss"] == out["loss"] and abs(out["loss"] - ref_out["loss"]) < 2e-3 * max(1.0, abs(ref_out["loss"])), \
        (out, ref_out)
    assert abs(val["log_loss"] - ref_val["log_loss"]) < 2e-3 * max(1.0, abs(ref_val["log_loss"])) and \
        abs(val["auc"] - ref_val["auc"]) < 5e-3, (val, ref_val)


def test_unmodified_pep_script_loop_with_clip_grad_on_installed_class(tmp_path):
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200 as R

    pep = refimport.load_script("scripts/deepfm/train_deepfm_pep.py", "ref_train_deepfm_pep")
    import src.models as ref_models

    emb = {"name": "pep", "threshold_type": "feature_dim", "init_threshold": -4.0}
    torch.manual_seed(2023)
    ref_model = ref_models.get_ctr_model(KDD, _cfg(dict(emb, checkpoint_weight_dir=str(tmp_path / "ref")), 0.0))
    with torch.no_grad():
        ref_model.embedding.emb.weight.uniform_(-0.1, 0.1)
    state = copy.deepcopy(ref_model.state_dict())
    loader = _loader(KDD)
    ref_opt = torch.optim.Adam(ref_model.parameters(), lr=1e-3, weight_decay=1e-5)       # train_deepfm_pep.py:185-189
    ref_out = pep.train_epoch(loader, ref_model, ref_opt, DEV, log_step=10, clip_grad=100)
    ref_sp = ref_model.embedding.get_sparsity(True)

    with refimport.installed() as mods:
        model = mods["src.models"].get_ctr_model(KDD, _cfg(dict(emb, checkpoint_weight_dir=str(tmp_path / "ours")), 0.0))
        assert isinstance(model.embedding, R.PepEmbeeding)
        model.load_state_dict(state, strict=True)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
        out = pep.train_epoch(loader, model, opt, DEV, log_step=10, clip_grad=100)      # the script's own loop
        sp = model.embedding.get_sparsity(True)
    assert abs(out["loss"] - ref_out["loss"]) < 2e-3 * max(1.0, abs(ref_out["loss"])), (out, ref_out)
    assert abs(sp[0] - ref_sp[0]) < 1e-4 and abs(sp[1] - ref_sp[1]) <= 1e-4 * ref_sp[1] + 16, (sp, ref_sp)


def test_unmodified_optembed_script_loop_on_installed_class():
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200 as R

    opt_script = refimport.load_script("scripts/deepfm/train_deepfm_optembed.py", "ref_train_deepfm_optembed")
    import src.models as ref_models

    def optimizers(model):               # train_deepfm_optembed.py:217-243, verbatim grouping
        groups = {"t_param": [], "default": []}
        for name, p in model.named_parameters():
            groups["t_param" if "t_param" in name else "default"].append(p)
        return [torch.optim.Adam([{"params": groups["default"], "lr": 1e-3, "weight_decay": 1e-5}]),
                torch.optim.SGD([{"params": groups["t_param"], "lr": 1e-4, "weight_decay": 0}])]

    emb = {"name": "deepfm_optembed", "t_init": 0.05}
    torch.manual_seed(2023)
    ref_model = ref_models.get_ctr_model(KDD, _cfg(emb))
    with torch.no_grad():
        ref_model.embedding._weight.uniform_(-0.01, 0.01)         # ||e||_1 ~ 0.08: around the thresholds
    state = copy.deepcopy(ref_model.state_dict())
    loader = _loader(KDD)
    torch.manual_seed(1)
    ref_out = opt_script.train_epoch(loader, ref_model, optimizers(ref_model), DEV, log_step=10, alpha=1e-4)
    with refimport.installed() as mods:
        # the real flow: install first, then run the unchanged script - its `from ...deepfm_opt_embed import OptEmbed`
        # (used for an isinstance check in the loop) then binds our class
        opt_script = refimport.load_script("scripts/deepfm/train_deepfm_optembed.py", "ref_train_deepfm_optembed2")
        assert opt_script.OptEmbed is R.OptEmbed
        model = mods["src.models"].get_ctr_model(KDD, _cfg(emb))
        assert isinstance(model.embedding, R.OptEmbed)
        model.load_state_dict(state, strict=True)
        torch.manual_seed(1)
        out = opt_script.train_epoch(loader, model, optimizers(model), DEV, log_step=10, alpha=1e-4)
    for k in ("loss", "loss_s"):
        assert abs(out[k] - ref_out[k]) < 2e-3 * max(1.0, abs(ref_out[k])), (k, out, ref_out)
    assert abs(out["sparsity"] - ref_out["sparsity"]) < 1e-3, (out, ref_out)
