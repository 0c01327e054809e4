"""The CUDA path against the numpy oracle (oracle/ctr_oracle.py) AT the BASELINE.json shapes, through the public
plugin API (`embedding.lookup`, the fused gather + first order + FM and its three-stage backward):

  Criteo F=39 / D=16 / B=2048  vanilla and QR (divider 2 / 5 / 20, mult): emb, y_fm, per-lookup g_emb, table / fc /
                               bias gradients, whole-model eval logits, one SparseAdam step of the fused row update
  KDD    F=11 / B=8192         PEP (feature_dim, ~80 % pruned) and OptEmbed (norm 1, supplied mask-D draw)
  Avazu  F=22 / B=2048         DCN-Mix cross head forward / backward and whole-model eval logits

These instantiate the production kernel templates (`<kind, V=4, LPR=4>`, D = 16), which the toy-sized golden files
(D = 8) do not.  Index / mask work is compared bit for bit; fp32 within rtol 1e-5 + atol 1e-5*max|ref|."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
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


def _ids(dims, b, seed, zipf=False):
    rng = np.random.default_rng(seed)
    if zipf:
        cols = [np.minimum(rng.zipf(1.05, b) - 1, d - 1) for d in dims]
    else:
        cols = [rng.integers(0, d, b) for d in dims]
    return np.stack(cols, 1).astype(np.int64)


def _np(t):
    return t.detach().cpu().numpy()


def _state_np(model):
    return {k: _np(v) for k, v in model.state_dict().items() if torch.is_tensor(v)}


def _lookup_grads(model, x_np, g_y, g_deep, params):
    """Run the fused lookup + its backward with the given upstream gradients; returns (emb, y_fm, grads)."""
    x = torch.from_numpy(x_np).to(DEV)
    emb, y = model.embedding.lookup(x, model.offsets, model.fc.weight, model._bias)
    grads = torch.autograd.grad([emb, y], params, [torch.from_numpy(g_deep).to(DEV), torch.from_numpy(g_y).to(DEV)],
                                allow_unused=True)
    return emb.detach(), y.detach(), grads


@pytest.mark.parametrize("emb_cfg", [{"name": "vanilla"}, {"name": "qr", "divider": 2}, {"name": "qr", "divider": 5},
                                     {"name": "qr", "divider": 20}], ids=lambda c: c["name"] + str(c.get("divider", "")))
@pytest.mark.parametrize("zipf", [False, True], ids=["uniform", "zipf"])
def test_criteo_shape_lookup_and_backward_match_the_oracle(R, emb_cfg, zipf):
    torch.manual_seed(3)
    b, dims = 2048, CRITEO_DIMS
    model = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=0.0, use_batchnorm=True,
                                       embedding_config=dict(emb_cfg))).to(DEV)
    with torch.no_grad():
        model.fc.weight.uniform_(-0.05, 0.05)
        model._bias.fill_(0.1)
        if emb_cfg["name"] == "qr":      # the reference's uniform(sqrt(1/N), 1) init makes logits O(800): tame it
            model.embedding.emb1.weight.mul_(0.5)
            model.embedding.emb2.weight.mul_(0.2)
    st = _state_np(model)
    x = _ids(dims, b, 5, zipf)
    rows = O.add_offsets(x, O.field_offsets(dims))
    rng = np.random.default_rng(9)
    g_y = (rng.standard_normal(b) / b).astype(np.float32)
    g_deep = (rng.standard_normal((b, len(dims), 16)) / b).astype(np.float32)
    params = list(model.embedding.parameters()) + [model.fc.weight, model._bias]
    emb, y, grads = _lookup_grads(model, x, g_y, g_deep, params)

    # ---- forward -----------------------------------------------------------------------------------------
    if emb_cfg["name"] == "vanilla":
        w = st["embedding._emb_module.weight"]
        emb_ref = O.gather_rows(w, rows)
    else:
        e1, e2 = st["embedding.emb1.weight"], st["embedding.emb2.weight"]
        i1, i2 = O.qr_indices(rows, emb_cfg["divider"])
        assert e1.shape[0] == emb_cfg["divider"] and e2.shape[0] == (sum(dims) - 1) // emb_cfg["divider"] + 1
        assert int(i2.max()) < e2.shape[0] and int(i1.max()) < e1.shape[0]
        emb_ref = O.qr_forward(e1, e2, rows, emb_cfg["divider"], "mult")
    np.testing.assert_array_equal(_np(emb), emb_ref)       # gather / one fp32 multiply: bit exact
    y_ref = O.deepfm_yfm(emb_ref.astype(np.float64), st["fc.weight"].astype(np.float64), st["_bias"].astype(np.float64),
                         rows)[:, 0]
    assert_close(_np(y), y_ref, what="y_fm")
    # whole model, eval mode (BatchNorm running stats), int32 ids == int64 ids
    model.eval()
    with torch.no_grad():
        logits = model(torch.from_numpy(x).to(DEV))
        assert torch.equal(logits, model(torch.from_numpy(x.astype(np.int32)).to(DEV)))
    st64 = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in st.items()}
    assert_close(_np(logits), O.deepfm_logits_eval(st64, emb_ref.astype(np.float64), rows), what="eval logits")
    model.train()

    # ---- backward ----------------------------------------------------------------------------------------
    g_emb = O.fm_backward(emb_ref.astype(np.float64), g_y.astype(np.float64), g_deep.astype(np.float64))
    g_fc, g_bias = O.first_order_backward(rows, g_y.astype(np.float64), sum(dims))
    if emb_cfg["name"] == "vanilla":
        assert_close(_np(grads[0]), O.scatter_add_dense(rows, g_emb, sum(dims)), what="table grad")
        touched = np.zeros(sum(dims), bool)
        touched[rows.reshape(-1)] = True
        assert float(np.abs(_np(grads[0])[~touched]).sum()) == 0.0
    else:
        g1, g2 = O.qr_backward(e1.astype(np.float64), e2.astype(np.float64), rows, emb_cfg["divider"], "mult", g_emb)
        assert_close(_np(grads[0]), g1, what="emb1 grad", atol_scale=2e-5)
        assert_close(_np(grads[1]), g2, what="emb2 grad")
    assert_close(_np(grads[-2]), g_fc, what="fc grad", atol_scale=2e-5)
    assert_close(_np(grads[-1]), g_bias, what="bias grad", atol_scale=2e-5)


def test_criteo_shape_per_lookup_gradient_and_fused_sparse_adam_step(R):
    """nn.Embedding(sparse=True) semantics: the COO gradient's values ARE the per-lookup g_emb; then the same
    gradient through the fused segmented-reduce + SparseAdam row update against the oracle's SparseAdam."""
    torch.manual_seed(4)
    b, dims = 2048, CRITEO_DIMS
    model = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=0.0, use_batchnorm=True,
                                       embedding_config={"name": "vanilla", "sparse": True})).to(DEV)
    with torch.no_grad():
        model.fc.weight.uniform_(-0.05, 0.05)
    w0 = _np(model.embedding.get_weight()).copy()
    x = _ids(dims, b, 6)
    rows = O.add_offsets(x, O.field_offsets(dims))
    rng = np.random.default_rng(10)
    g_y = (rng.standard_normal(b) / b).astype(np.float32)
    g_deep = (rng.standard_normal((b, len(dims), 16)) / b).astype(np.float32)
    table = model.embedding.get_weight()
    emb, y, (g_table,) = _lookup_grads(model, x, g_y, g_deep, [table])
    assert g_table.is_sparse and g_table._nnz() == b * len(dims)
    np.testing.assert_array_equal(_np(g_table._indices())[0], rows.reshape(-1))
    g_emb = O.fm_backward(w0[rows].astype(np.float64), g_y.astype(np.float64), g_deep.astype(np.float64))
    assert_close(_np(g_table._values()), g_emb.reshape(-1, 16), what="per-lookup g_emb")

    opt = R.FusedSparseAdam(model.embedding, lr=1e-3)
    wo, mo, vo = w0.copy(), np.zeros_like(w0), np.zeros_like(w0)
    for step in (1, 2):
        xs = _ids(dims, b, 20 + step, zipf=(step == 2))
        rs = O.add_offsets(xs, O.field_offsets(dims))
        opt.zero_grad()
        e, yy = model.embedding.lookup(torch.from_numpy(xs).to(DEV), model.offsets, model.fc.weight, model._bias)
        torch.autograd.backward([e, yy], [torch.from_numpy(g_deep).to(DEV), torch.from_numpy(g_y).to(DEV)])
        assert table.grad is None                      # consumed inside the backward (opt-in fused mode)
        opt.step()
        ge = O.fm_backward(wo[rs].astype(np.float64), g_y.astype(np.float64), g_deep.astype(np.float64))
        uniq, sums = O.coalesce_rows(rs, ge.astype(np.float32).reshape(-1, 16))
        O.sparse_adam_rows(wo, mo, vo, step, uniq, sums, lr=1e-3)
        assert_close(_np(table), wo, what=f"table after step {step}", atol_scale=2e-5)
        st = opt.state[table]
        assert_close(_np(st["exp_avg"]), mo, what="exp_avg", atol_scale=2e-5)
        assert_close(_np(st["exp_avg_sq"]), vo, what="exp_avg_sq", atol_scale=2e-5)
    opt.detach_from_module()


@pytest.mark.parametrize("tt", ["feature_dim", "feature"])
def test_kdd_shape_pep_matches_the_oracle_with_bit_exact_masks(R, tt, tmp_path):
    import recsys_benchmark_b200.functional as RF

    torch.manual_seed(5)
    b, dims = 8192, KDD_DIMS
    model = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=0.0, use_batchnorm=True,
                                       embedding_config={"name": "pep", "threshold_type": tt,
                                                         "checkpoint_weight_dir": str(tmp_path)})).to(DEV)
    emb_mod = model.embedding
    g = torch.Generator(device=DEV).manual_seed(1)
    with torch.no_grad():
        emb_mod.emb.weight.uniform_(-0.5, 0.5, generator=g)
        emb_mod.s.copy_(-0.4 + 0.3 * torch.randn(emb_mod.s.shape, generator=g, device=DEV))
        model.fc.weight.uniform_(-0.05, 0.05)
    # the threshold sigmoid(s) in the reference's arithmetic = torch.sigmoid on the device it trains on; the kernels'
    # own sigmoid must be that function bit for bit (then every |v| > sigmoid(s) decision is the reference's)
    sig_t = torch.sigmoid(emb_mod.s.detach())
    assert torch.equal(RF.sigmoid(emb_mod.s.detach()), sig_t)
    w, s, sig = _np(emb_mod.emb.weight), _np(emb_mod.s), _np(sig_t)
    x = _ids(dims, b, 7)
    rows = O.add_offsets(x, O.field_offsets(dims))
    uniq = np.unique(rows)
    remap = np.searchsorted(uniq, rows)                  # the oracle works on the touched rows only (6 M-row table)
    ws, ss, sigs = w[uniq], (s[uniq] if s.shape[0] == w.shape[0] else s), (sig[uniq] if s.shape[0] == w.shape[0] else sig)
    rng = np.random.default_rng(11)
    g_y = (rng.standard_normal(b) / b).astype(np.float32)
    g_deep = (rng.standard_normal((b, len(dims), 16)) / b).astype(np.float32)
    emb, y, grads = _lookup_grads(model, x, g_y, g_deep, [emb_mod.emb.weight, emb_mod.s, model.fc.weight])
    emb_ref = O.pep_forward(ws, ss, remap, sig=sigs)
    np.testing.assert_array_equal(_np(emb), emb_ref)     # bit exact, masks included
    pruned = float((emb_ref == 0).mean())
    assert 0.6 < pruned < 0.95, pruned
    g_emb = O.fm_backward(emb_ref.astype(np.float64), g_y.astype(np.float64), g_deep.astype(np.float64))
    g_w, g_s = O.pep_backward(ws.astype(np.float64), ss.astype(np.float64), remap, g_emb, sig=sigs.astype(np.float64))
    got_w, got_s = _np(grads[0]), _np(grads[1])
    assert_close(got_w[uniq], g_w, what="weight grad")
    assert_close(got_s[uniq] if s.shape[0] == w.shape[0] else got_s, g_s, what="s grad", atol_scale=2e-5)
    mask = np.ones(w.shape[0], bool)
    mask[uniq] = False
    assert float(np.abs(got_w[mask]).sum()) == 0.0
    # full-table bookkeeping: get_sparsity counts exactly the reference's non-zeros
    sp, nnz = emb_mod.get_sparsity(True)
    ref_nnz = int(torch.count_nonzero(torch.sign(emb_mod.emb.weight) * torch.relu(emb_mod.emb.weight.abs() - sig_t)))
    assert nnz == ref_nnz and abs(sp - (1 - ref_nnz / w.size)) < 1e-12


def test_kernel_sigmoid_is_torch_cuda_sigmoid_bit_for_bit():
    import recsys_benchmark_b200.functional as RF

    g = torch.Generator(device=DEV).manual_seed(0)
    s = torch.cat([torch.randn(1 << 22, generator=g, device=DEV) * 4, torch.linspace(-200, 200, 100001, device=DEV),
                   torch.tensor([0.0, -0.0, -150.0, 88.7, -88.7, 103.9, -103.9, 1e-30, -1e-30, float("inf"),
                                 -float("inf")], device=DEV)])
    assert torch.equal(RF.sigmoid(s), torch.sigmoid(s))


@pytest.mark.parametrize("norm", [1, 2])
def test_kdd_shape_optembed_supernet_matches_the_oracle(R, norm):
    torch.manual_seed(6)
    b, dims = 8192, KDD_DIMS
    cfg = {"name": "deepfm_optembed"}
    if norm == 2:
        cfg["norm"] = 2
    model = R.get_ctr_model(dims, dict(num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=0.0, use_batchnorm=True,
                                       embedding_config=cfg)).to(DEV)
    emb_mod = model.embedding
    g = torch.Generator(device=DEV).manual_seed(2)
    with torch.no_grad():
        emb_mod._weight.uniform_(-0.2, 0.2, generator=g)
        t = emb_mod._mask_e_module._t_param
        lo, span = (1.2, 0.8) if norm == 1 else (0.40, 0.12)
        t.copy_(lo + span * torch.rand(t.shape, generator=g, device=DEV))
        model.fc.weight.uniform_(-0.05, 0.05)
    model.train()
    x = _ids(dims, b, 8)
    rows = O.add_offsets(x, O.field_offsets(dims))
    uniq = np.unique(rows)
    remap = np.searchsorted(uniq, rows)
    w = _np(emb_mod._weight)[uniq]
    t_np = _np(emb_mod._mask_e_module._t_param)
    # the reference draws k = torch.randint(0, D, (B, F), device) (deepfm_opt_embed.py:222-224): same call, same seed
    torch.manual_seed(77)
    k = _np(torch.randint(0, 16, (b, len(dims)), device=DEV))
    rng = np.random.default_rng(12)
    g_y = (rng.standard_normal(b) / b).astype(np.float32)
    g_deep = (rng.standard_normal((b, len(dims), 16)) / b).astype(np.float32)
    torch.manual_seed(77)
    emb, y, grads = _lookup_grads(model, x, g_y, g_deep, [emb_mod._weight, emb_mod._mask_e_module._t_param])
    emb_ref = O.optembed_train_forward(w, t_np, remap, k, norm=norm)
    got = _np(emb)
    # mask-D (integer index) is bit exact everywhere; mask-E may flip only where ||e|| - t is within rounding of zero
    differs = (got != emb_ref).any(-1)
    z = O._row_norm(w[remap].astype(np.float64), norm) - t_np[None, :].astype(np.float64)
    assert float(np.abs(z[differs]).max(initial=0.0)) < 2e-6 and int(differs.sum()) <= 2
    kept = float((np.abs(emb_ref).sum(-1) > 0).mean())
    assert 0.2 < kept < 0.9, kept
    if not differs.any():
        g_emb = O.fm_backward(emb_ref.astype(np.float64), g_y.astype(np.float64), g_deep.astype(np.float64))
        g_w, g_t = O.optembed_train_backward(w.astype(np.float64), t_np.astype(np.float64), remap, k, g_emb, norm=norm)
        assert_close(_np(grads[0])[uniq], g_w, what="weight grad", atol_scale=2e-5)
        assert_close(_np(grads[1]), g_t, what="t grad", atol_scale=5e-5)


def test_avazu_shape_dcn_mix_matches_the_oracle(R):
    torch.manual_seed(7)
    b, dims = 2048, AVAZU_DIMS
    model = R.get_ctr_model(dims, dict(name="dcn_mix", num_factor=16, hidden_sizes=[400, 400, 400], p_dropout=0.0,
                                       compile_model=False, embedding_config={"name": "vanilla"})).to(DEV)
    with torch.no_grad():
        for bia in model.cross_head.biases:
            bia.normal_(0, 0.1)
        model.embedding.get_weight().mul_(30.0)        # xavier over 1 M rows is ~2e-3: give the cross layers signal
    st = _state_np(model)
    x = _ids(dims, b, 9)
    rows = O.add_offsets(x, O.field_offsets(dims))
    head = model.cross_head
    nl = len(head.biases)
    params64 = dict(V=[_np(head.V[l]).astype(np.float64) for l in range(nl)],
                    C=[_np(head.C[l]).astype(np.float64) for l in range(nl)],
                    U=[_np(head.U[l]).astype(np.float64) for l in range(nl)],
                    biases=[_np(head.biases[l]).astype(np.float64) for l in range(nl)],
                    gates=_np(head.gates).astype(np.float64))
    x0 = O.gather_rows(st["embedding._emb_module.weight"], rows).reshape(b, -1).astype(np.float64)
    xl_ref, caches = O.dcn_mix_forward(x0, params64)
    # cross head alone, forward + backward
    x0_t = torch.from_numpy(x0.astype(np.float32)).to(DEV).requires_grad_(True)
    out = head(x0_t)
    assert_close(_np(out), xl_ref, what="cross head out", atol_scale=2e-5)
    g_out = (np.random.default_rng(13).standard_normal(xl_ref.shape) / b).astype(np.float32)
    plist = [x0_t] + [head.V[l] for l in range(nl)] + [head.C[l] for l in range(nl)] + [head.U[l] for l in range(nl)] \
        + [head.biases[l] for l in range(nl)] + [head.gates]
    grads = torch.autograd.grad(out, plist, torch.from_numpy(g_out).to(DEV))
    g_x0, gref = O.dcn_mix_backward(x0, params64, caches, g_out.astype(np.float64))
    assert_close(_np(grads[0]), g_x0, what="g_x0", atol_scale=2e-5)
    off = 1
    for name in ("V", "C", "U", "biases"):
        for l in range(nl):
            assert_close(_np(grads[off]), gref[name][l], what=f"g_{name}[{l}]", atol_scale=3e-5)
            off += 1
    assert_close(_np(grads[off]), gref["gates"], what="g_gates", atol_scale=3e-5)
    # whole model, eval mode
    model.eval()
    with torch.no_grad():
        logits = model(torch.from_numpy(x).to(DEV))
    st64 = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in st.items()}
    ref_logits = O.mlp_eval(xl_ref, O.mlp_layers_from_state(st64, "_dnn."))[:, 0]
    assert_close(_np(logits), ref_logits, what="eval logits", atol_scale=2e-5)
