"""Parity against the UNMODIFIED reference at the BASELINE.json shapes, on the same GPU.

The reference (staged under baseline/_ref by __graft_entry__.build(), see tests/refimport.py) builds its own
DeepFM / DCN_Mix with its own embedding plugins and runs them with torch's CUDA operators; our model is built
from the same config, loads the reference's state dict with strict=True, and must reproduce on identical
synthetic Criteo- / KDD- / Avazu-shaped batches

  * the plugin output `embedding(rows)`            - bit for bit where the arithmetic is a gather, a single
                                                     multiply or a threshold (vanilla, QR, PEP, masks),
  * the logits (eval and train mode)               - fp32 tolerance below,
  * every parameter gradient (table, QR tables, thresholds, first-order weights, MLP / cross layers),
  * the weights after one optimizer step of the reference's own `get_optimizers` recipe.

fp32 tolerance: rtol 1e-5 with atol = 1e-5 * max|ref|, OR - for quantities that go through the
GEMMs, where both sides are a rounding of the exact result - an error against the reference's fp64 twin no larger
than 4x the error of the reference's own fp32 run against that twin.
"""
import copy

import numpy as np
import pytest
import torch

from tests import refimport

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refimport.available(), reason="baseline/_ref not staged")]
DEV = "cuda:0"

CRITEO_DIMS = [49, 101, 126, 45, 223, 118, 84, 76, 95, 9, 30, 40, 75, 1458, 555, 193949, 138801, 306, 19, 11970, 634,
               4, 42646, 5178, 192773, 3175, 27, 11422, 181075, 11, 4654, 2032, 5, 189657, 18, 16, 59697, 86, 45571]
AVAZU_DIMS = [100000] * 10 + [1000] * 12
KDD_DIMS = [600000] * 8 + [400000] * 3


@pytest.fixture(scope="module")
def R():
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200 as r

    return r


@pytest.fixture(scope="module")
def REF():
    refimport.activate()
    import src.models as ref_models
    import src.models.deepfm as ref_deepfm

    # these tests need the reference's OWN classes: fail loudly if another test left the registry rebound
    assert ref_models.DeepFM.__module__.startswith("src."), "reference registry is rebound to our classes"
    return ref_models, ref_deepfm


def _batch(dims, b, seed, dtype=torch.int64):
    g = torch.Generator().manual_seed(seed)
    x = torch.stack([torch.randint(0, d, (b,), generator=g) for d in dims], 1).to(dtype)
    y = torch.randint(0, 2, (b,), generator=g)
    return x.to(DEV), y.to(DEV)


def _dense(g):
    return g.to_dense() if g.is_sparse else g


def _within(got, ref, rtol=1e-5, atol_scale=1e-5):
    got, ref = got.double(), ref.double()
    atol = atol_scale * float(ref.abs().max()) if ref.numel() else 0.0
    return bool(((got - ref).abs() <= atol + rtol * ref.abs()).all())


def assert_parity(got, ref32, ref64, what):
    """rtol 1e-5 / atol 1e-5*max|ref| against the reference's fp32 result, or no further from the reference's fp64
    twin than 4x the reference's own fp32 distance to it (both from SURVEY.md 8c)."""
    assert got.shape == ref32.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref32.shape)}"
    if _within(got, ref32):
        return
    scale = float(ref64.abs().max()) or 1.0
    d_ours = (got.double() - ref64).abs()
    e_ours = float(d_ours.max()) / scale
    e_ref = float((ref32.double() - ref64).abs().max()) / scale
    i = int(d_ours.argmax())
    bad = int(((got.double() - ref32.double()).abs() > 1e-5 * float(ref32.abs().max()) + 1e-5 * ref32.double().abs()).sum())
    assert e_ours <= 4.0 * e_ref + 1e-7, (
        f"{what}: max err vs fp64 twin {e_ours:.3e} (relative to max|ref|), reference fp32 itself {e_ref:.3e}; worst "
        f"flat index {i}: ours {float(got.reshape(-1)[i]):.9e} ref32 {float(ref32.reshape(-1)[i]):.9e} "
        f"ref64 {float(ref64.reshape(-1)[i]):.9e}; {bad}/{got.numel()} elements outside rtol/atol vs ref32")


def _find_node(fn, name, seen=None):
    """Walk the autograd graph from `fn` for the first node whose class name is `name`."""
    seen = set() if seen is None else seen
    if fn is None or fn in seen:
        return None
    seen.add(fn)
    if type(fn).__name__ == name:
        return fn
    for nxt, _ in fn.next_functions:
        hit = _find_node(nxt, name, seen)
        if hit is not None:
            return hit
    return None


class _ReluTies:
    """ReLU is discontinuous in its gradient: a pre-activation within fp32 rounding of zero can land on either side
    when the Linear before it is computed by two different (equally accurate) GEMMs, and BatchNorm's batch statistics
    then spread that one decision over every sample's gradient.  The reference's BatchNorm1d outputs (= ReLU inputs)
    are recorded with forward hooks; ours come from the fused dense-tail node (its keep-and-positive masks: dropout is
    off in these tests, so mask = [pre-activation > 0]) or, on the layer-by-layer path, from the same hooks.
    `flips()` returns how many decisions differ and checks that each of them really is a tie (|pre-activation| < 2e-5
    after normalisation)."""

    def __init__(self, ref, ours):
        self.pre = {"ref": [], "ours": []}
        self.handles = []
        for name, m in (("ref", ref), ("ours", ours)):
            for mod in m.modules():
                if isinstance(mod, torch.nn.BatchNorm1d):
                    self.handles.append(mod.register_forward_hook(
                        lambda _m, _i, out, _n=name: self.pre[_n].append(out.detach() > 0)))
        self.ref_values = []
        for mod in ref.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                self.handles.append(mod.register_forward_hook(lambda _m, _i, out: self.ref_values.append(out.detach())))

    def flips(self, ours_logits=None):
        if not self.pre["ours"] and ours_logits is not None:
            node = _find_node(ours_logits.grad_fn, "_MlpBatchNormBackward")
            if node is not None:
                self.pre["ours"] = [m > 0 for m in node.masks]
        n = 0
        assert len(self.pre["ref"]) == len(self.pre["ours"]), (len(self.pre["ref"]), len(self.pre["ours"]))
        for a, b, v in zip(self.pre["ref"], self.pre["ours"], self.ref_values):
            diff = a != b
            if bool(diff.any()):
                assert float(v[diff].abs().max()) < 2e-5, "a ReLU decision differs away from a rounding tie"
                n += int(diff.sum())
        return n

    def remove(self):
        for h in self.handles:
            h.remove()


def assert_parity_given_ties(got, ref32, ref64, what, flips):
    """No ReLU tie flipped: the strict rule.  Otherwise (see _ReluTies) the two gradients are those of two
    neighbouring branches of a piecewise-linear function: close in norm, not element-wise."""
    if flips == 0:
        return assert_parity(got, ref32, ref64, what)
    if _within(got, ref32):
        return
    rel = float((got.double() - ref32.double()).norm() / ref32.double().norm().clamp_min(1e-300))
    assert rel < 5e-2 * flips, f"{what}: relative L2 difference {rel:.3e} with {flips} ReLU tie flip(s)"


def _noise_keys(model):
    """Linear biases that feed a BatchNorm: their gradient is zero by construction, both sides hold fp32 noise."""
    out = set()
    for name in ("_deep_branch", "_dnn"):
        seq = getattr(model, name, None)
        if seq is None:
            continue
        mods = list(seq)
        for i, m in enumerate(mods[:-1]):
            if isinstance(m, torch.nn.Linear) and isinstance(mods[i + 1], torch.nn.BatchNorm1d):
                out.add(f"{name}.{i}.bias")
    return out


def _tweak(model, kind):
    """Deterministic non-trivial parameter values for the pruning variants (xavier weights of a 6 M-row table are
    ~1e-3, far below any threshold: everything would be pruned and the test vacuous)."""
    g = torch.Generator(device=DEV).manual_seed(99)
    emb = model.embedding
    with torch.no_grad():
        if kind == "pep":
            emb.emb.weight.uniform_(-0.5, 0.5, generator=g)
            emb.s.copy_(-0.4 + 0.3 * torch.randn(emb.s.shape, generator=g, device=DEV))    # ~80 % pruned
        elif kind == "qr":
            # the reference's uniform(sqrt(1/N), 1) init (qr_embedding.py:74-84) makes fresh logits O(800): the loss
            # saturates and most gradients are rounding noise.  Scale the tables so that the comparison says something.
            emb.emb1.weight.mul_(0.5)
            emb.emb2.weight.mul_(0.2)
        elif kind == "optembed":
            emb._weight.uniform_(-0.2, 0.2, generator=g)
            t = emb._mask_e_module._t_param
            t.copy_(1.2 + 0.8 * torch.rand(t.shape, generator=g, device=DEV))
        elif kind == "optembed_d":
            emb._weight.uniform_(-0.2, 0.2, generator=g)
        model.fc.weight.uniform_(-0.05, 0.05, generator=g) if hasattr(model, "fc") else None


CASES = {
    # name: (dims, batch, model cfg, optimizer cfg, tweak, exact plugin output?)
    "criteo_vanilla_adam": (CRITEO_DIMS, 2048, dict(embedding_config={"name": "vanilla"}),
                            dict(learning_rate=1e-3, weight_decay=1e-6), None, True),
    "criteo_vanilla_sparse_adam": (CRITEO_DIMS, 2048, dict(embedding_config={"name": "vanilla", "sparse": True}),
                                   dict(learning_rate=1e-3, weight_decay=1e-6, sparse=True), None, True),
    "criteo_qr2": (CRITEO_DIMS, 2048, dict(embedding_config={"name": "qr", "divider": 2}),
                   dict(learning_rate=1e-3, weight_decay=1e-6), "qr", True),
    "criteo_qr5": (CRITEO_DIMS, 2048, dict(embedding_config={"name": "qr", "divider": 5}),
                   dict(learning_rate=1e-3, weight_decay=1e-6), "qr", True),
    "criteo_qr20": (CRITEO_DIMS, 2048, dict(embedding_config={"name": "qr", "divider": 20}),
                    dict(learning_rate=1e-3, weight_decay=1e-6), "qr", True),
    "criteo_qr_default_add": (CRITEO_DIMS, 2048, dict(embedding_config={"name": "qr", "operation": "add"}),
                              dict(learning_rate=1e-3, weight_decay=1e-6), "qr", True),
    "criteo_qr5_cat": (CRITEO_DIMS, 2048, dict(embedding_config={"name": "qr", "divider": 5, "operation": "cat"}),
                       dict(learning_rate=1e-3, weight_decay=1e-6), "qr", True),
    "kdd_pep_feature_dim": (KDD_DIMS, 8192, dict(embedding_config={"name": "pep", "threshold_type": "feature_dim"}),
                            dict(learning_rate=1e-3, weight_decay=1e-5), "pep", True),
    "kdd_pep_feature": (KDD_DIMS, 8192, dict(embedding_config={"name": "pep", "threshold_type": "feature"}),
                        dict(learning_rate=1e-3, weight_decay=1e-5), "pep", True),
    "kdd_optembed": (KDD_DIMS, 8192, dict(embedding_config={"name": "deepfm_optembed"}),
                     dict(learning_rate=3e-5, weight_decay=1e-3), "optembed", False),
    "kdd_optembed_d": (KDD_DIMS, 8192, dict(embedding_config={"name": "deepfm_optembed_d"}),
                       dict(learning_rate=3e-5, weight_decay=1e-3), "optembed_d", True),
    "avazu_dcn_mix": (AVAZU_DIMS, 2048, dict(name="dcn_mix", compile_model=False,
                                             embedding_config={"name": "vanilla"}),
                      dict(learning_rate=1e-3, weight_decay=1e-6), None, True),
}


def _build(mod, dims, cfg, tmp_path):
    cfg = copy.deepcopy(cfg)
    cfg.setdefault("num_factor", 16)
    cfg.setdefault("hidden_sizes", [400, 400, 400])
    cfg.setdefault("p_dropout", 0.0)           # dropout off: the two sides draw different masks by design
    if cfg.get("name") != "dcn_mix":
        cfg.setdefault("use_batchnorm", True)
    if cfg["embedding_config"]["name"].startswith("pep"):
        cfg["embedding_config"]["checkpoint_weight_dir"] = str(tmp_path)
    return mod.get_ctr_model(dims, cfg)


@pytest.mark.parametrize("case", sorted(CASES))
def test_model_matches_the_unmodified_reference_on_the_same_gpu(R, REF, case, tmp_path):
    ref_models, ref_deepfm = REF
    dims, b, cfg, opt_cfg, tweak, exact_emb = CASES[case]
    torch.manual_seed(2023)
    ref = _build(ref_models, dims, cfg, tmp_path / "ref").to(DEV)
    assert type(ref).__module__.startswith("src."), "must be the reference's own class"
    if tweak:
        _tweak(ref, tweak)
    ours = _build(R, dims, cfg, tmp_path / "ours")
    ours.load_state_dict(ref.state_dict(), strict=True)      # frozen names / shapes / dtypes
    ours.to(DEV)
    ref64 = copy.deepcopy(ref).double()
    is_opt = "optembed" in cfg["embedding_config"]["name"]
    x, y = _batch(dims, b, 7)
    rows = x + ref.offsets

    # ---- plugin forward: embedding(rows) ------------------------------------------------------------------
    for m in (ref, ours, ref64):
        m.train()
    outs = []
    for m in (ref, ours):
        torch.manual_seed(11)                 # OptEmbed draws its mask-D ids with torch.randint on the device
        outs.append(m.embedding(rows).detach())
    assert outs[0].shape == outs[1].shape
    if exact_emb:
        assert torch.equal(outs[0], outs[1]), f"{case}: plugin output differs from the reference's"
    else:
        # OptEmbed mask-E: ||e||_1 is summed in another order, so a row whose norm is within rounding of its
        # threshold may flip; everything else is exact
        diff_rows = (outs[0] != outs[1]).any(-1)
        if bool(diff_rows.any()):
            e = torch.nn.functional.embedding(rows, ref.embedding._weight).double()
            t = ref.embedding._mask_e_module._t_param.double()
            margin = (e.abs().sum(-1) - t[None, :]).abs()
            assert float(margin[diff_rows].max()) < 1e-5, "a mask differs away from a rounding tie"
            assert int(diff_rows.sum()) <= 2

    # ---- eval logits -----------------------------------------------------------------------------------------
    for m in (ref, ours, ref64):
        m.eval()
    with torch.no_grad():
        if is_opt:
            for m in (ref, ours, ref64):
                m.embedding.get_weight()      # fresh OptEmbed in eval mode needs its cache built first
        assert_parity(ours(x), ref(x), ref64(x), f"{case} eval logits")
        assert torch.equal(ours(x.int()), ours(x)), "int32 ids must give bit-identical logits"

    # ---- one training step of the reference's recipe -------------------------------------------------------------
    crit = torch.nn.BCEWithLogitsLoss()
    models = {"ref": ref, "ours": ours, "ref64": ref64}
    opts = {"ref": ref_deepfm.get_optimizers(ref, dict(opt_cfg)), "ours": R.get_optimizers(ours, dict(opt_cfg)),
            "ref64": []}
    logits, grads = {}, {}
    ours_masks = None
    ties = _ReluTies(ref, ours)
    for k, m in models.items():
        m.train()
        torch.manual_seed(12)
        out = m(x)
        loss = crit(out, y.to(out.dtype))
        for o in opts[k]:
            o.zero_grad()
        if k == "ours":
            ours_node = _find_node(out.grad_fn, "_MlpBatchNormBackward")
            ours_masks = [mm > 0 for mm in ours_node.masks] if ours_node is not None and ours_node.masks else None
        loss.backward()
        logits[k] = out.detach()
        grads[k] = {n: _dense(p.grad).detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    if ours_masks is not None and not ties.pre["ours"]:
        ties.pre["ours"] = ours_masks
    flips = ties.flips()
    ties.remove()
    assert flips <= 6, f"{flips} ReLU decisions differ"
    assert_parity(logits["ours"], logits["ref"], logits["ref64"], f"{case} train logits")
    noise = _noise_keys(ref)
    assert set(grads["ours"]) == set(grads["ref"]), "the same parameters must receive gradients"
    for n in grads["ref"]:
        if n in noise:
            assert float(grads["ours"][n].abs().max()) < 1e-5
            continue
        assert_parity_given_ties(grads["ours"][n], grads["ref"][n], grads["ref64"][n], f"{case} grad {n}", flips)
    if not is_opt and "pep" not in case and "qr" not in case:
        # rows of the big table that no lookup touched have exactly zero gradient, on both sides
        big = max(grads["ref"], key=lambda n: grads["ref"][n].numel())
        touched = torch.zeros(grads["ref"][big].shape[0], dtype=torch.bool, device=DEV)
        touched[rows.reshape(-1)] = True
        assert float(grads["ours"][big][~touched].abs().sum()) == 0.0 == float(grads["ref"][big][~touched].abs().sum())
    for k in ("ref", "ours"):
        for o in opts[k]:
            o.step()
    lr = opt_cfg["learning_rate"]
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    for n, v in sd_ref.items():
        if not torch.is_floating_point(v) or n in noise:
            assert n in noise or torch.equal(v, sd_ours[n]), n
            continue
        # Both sides run the SAME torch optimizer on gradients that were just shown to agree.  Adam's first step is
        # lr * g / (|g| + eps): it moves every weight by <= lr whatever the gradient's size, so an element whose
        # gradient is at rounding level (|g| ~ 1e-7 against max|g| ~ 1e-2: inside the gradient tolerance, above
        # eps = 1e-8) may legitimately move +lr on one side and -lr on the other.  Hence: a hard bound of 2 lr, and
        # only a small share of the elements may be off by more than the fp32 tolerance at all.
        d = (v - sd_ours[n]).abs()
        assert float(d.max()) <= 2.02 * lr * (1 + 1e-3) + 1e-5 * float(v.abs().max()), \
            f"{case} after step {n}: max diff {float(d.max()):.3e}"
        tight = 1e-5 * float(v.abs().max()) + 1e-5 * v.abs() + 1e-3 * lr
        share = float((d > tight).float().mean())
        assert share < 2e-2 or share * v.numel() <= 8, \
            f"{case} after step {n}: {share:.2e} of the elements differ by more than the tolerance"


@pytest.mark.parametrize("fused", ["adam", "sgd"])
def test_fused_sparse_row_update_matches_the_reference_sparse_optimizers(R, REF, fused, tmp_path):
    """configs/deepfm/base_config_sparse.yaml recipe at the Criteo shape, three steps: the reference's SparseAdam /
    sparse SGD on its COO gradient vs our fused segmented-reduce + row update (`fused_sparse: true`).
    Each step first checks that our backward hands the row optimizer the reference's gradient (same lookups, values
    within fp32 tolerance), then feeds BOTH optimizers the reference's values, so that the comparison of the updated
    tables isolates the row-update arithmetic (Adam's g / (sqrt(v) + eps) amplifies rounding-level differences of
    near-zero gradient elements to the size of lr, which would say nothing about the update itself)."""
    ref_models, ref_deepfm = REF
    torch.manual_seed(5)
    cfg = dict(embedding_config={"name": "vanilla", "sparse": True})
    ref = _build(ref_models, CRITEO_DIMS, cfg, tmp_path).to(DEV)
    ours = _build(R, CRITEO_DIMS, cfg, tmp_path)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.to(DEV)
    opt_cfg = dict(learning_rate=1e-3, weight_decay=1e-6, sparse=True, optimizer=fused)
    o_ref = ref_deepfm.get_optimizers(ref, dict(opt_cfg))
    o_ours = R.get_optimizers(ours, dict(opt_cfg, fused_sparse=True))
    fused_opt = o_ours[0]
    crit = torch.nn.BCEWithLogitsLoss()
    for step in range(3):
        x, y = _batch(CRITEO_DIMS, 2048, 100 + step)
        # same parameters on both sides at the start of every step (the dense Adam of the MLP amplifies
        # rounding-level gradient differences to the size of lr; this test is about the table's row update)
        ours.load_state_dict(ref.state_dict(), strict=True)
        for m, opts in ((ref, o_ref), (ours, o_ours)):
            m.train()
            loss = crit(m(x), y.float())
            for o in opts:
                o.zero_grad()
            loss.backward()
        g_ref = ref.embedding.get_weight().grad
        (table, rows, rg, pair), = fused_opt._pending
        assert torch.equal(rows.reshape(-1), g_ref._indices()[0])
        # same lookups, same gradient up to fp32 rounding and ReLU ties (see _ReluTies; the element-wise rule is
        # applied to the COO gradient in test_model_matches_the_unmodified_reference_on_the_same_gpu)
        rel = float((rg.reshape(-1, 16) - g_ref._values()).norm() / g_ref._values().norm())
        assert rel < 5e-2, f"step {step}: per-lookup gradient differs from the reference's COO values ({rel:.3e})"
        rg.copy_(g_ref._values().reshape(rg.shape))
        for opts in (o_ref, o_ours):
            for o in opts:
                o.step()
        w_ref, w_ours = ref.embedding.get_weight(), ours.embedding.get_weight()
        d = (w_ref - w_ours).abs()
        assert float(d.max()) <= 2e-6 * float(w_ref.abs().max()) + 2e-3 * 1e-3, f"step {step}: {float(d.max()):.3e}"
        touched = torch.zeros(sum(CRITEO_DIMS), dtype=torch.bool, device=DEV)
        touched[(x + ref.offsets).reshape(-1)] = True
        assert torch.equal(w_ref[~touched], w_ours[~touched])
