"""GPU tests of the fp32-accurate tensor-core GEMM (rsb_gemm_planes: hand-written tcgen05 kernel on bf16 planes)
through the C ABI wrappers: every layout, batching, fused epilogue, and fp32-level accuracy
(error vs an fp64 product must be of the same order as cuBLAS fp32 SGEMM's)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def LA():
    import __graft_entry__ as G

    G.build()
    import recsys_benchmark_b200.linalg as la

    return la


def _err(got, ref64):
    return float((got.double() - ref64).abs().max() / ref64.abs().max())


@pytest.mark.parametrize("m,n,k", [(128, 128, 16), (256, 400, 624), (2048, 400, 400), (1000, 64, 352), (4096, 256, 352),
                                   (520, 352, 256), (4, 4, 4), (132, 12, 20),
                                   # KDD-shaped MLP (F = 11: 176-wide input) forward / dX / dW shapes
                                   (8192, 400, 176), (8192, 176, 400), (400, 176, 8192), (176, 400, 8192)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False)])
def test_gemm_layouts_match_fp64(LA, m, n, k, ta, tb):
    torch.manual_seed(m + n + k)
    a = torch.randn((k, m) if ta else (m, k), device=DEV)
    b = torch.randn((n, k) if tb else (k, n), device=DEV)
    bias = torch.randn(n, device=DEV)
    out = LA.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias)
    ref = (a.double().t() if ta else a.double()) @ (b.double().t() if tb else b.double()) + bias.double()
    cublas = (a.t() if ta else a) @ (b.t() if tb else b) + bias
    e_ours, e_cublas = _err(out, ref), _err(cublas, ref)
    assert out.shape == (m, n)
    assert e_ours < 2e-6, f"relative error {e_ours:.2e} (cuBLAS fp32: {e_cublas:.2e})"
    assert e_ours < max(4 * e_cublas, 5e-7)


@pytest.mark.parametrize("m,n,k", [(128, 128, 16), (256, 400, 624), (2048, 400, 400), (1000, 64, 352), (520, 352, 256),
                                   (132, 12, 20), (4, 4, 4)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("mag_a,mag_b", [(1.0, 1.0), (3e-7, 0.05), (2e4, 1e-3)])
def test_gemm_fp16x2_planes_match_fp64(LA, m, n, k, a_mn, b_mn, mag_a, mag_b):
    """Two scaled fp16 planes per operand, 3 MMAs per product (RSB_PLANES_FP16X2): the scale comes from a device-side
    max |x|, so tiny gradients (3e-7) and large activations (2e4) keep the same relative accuracy; the error against
    fp64 stays at the level of an fp32 GEMM's own rounding."""
    from recsys_benchmark_b200 import planes as P

    if (a_mn and m % 8) or (b_mn and n % 8):
        pytest.skip("MN-major operands need 8-aligned rows")
    torch.manual_seed(m + n + k)
    a = torch.randn((k, m) if a_mn else (m, k), device=DEV) * mag_a
    b = torch.randn((k, n) if b_mn else (n, k), device=DEV) * mag_b
    bias = torch.randn(n, device=DEV) * mag_a * mag_b
    pa, pb = P.split(a, fmt=P.FP16X2), P.split(b, fmt=P.FP16X2)
    assert pa.data.dtype == torch.float16 and pa.data.shape[0] == 2
    # the planes represent the operand to ~2^-22 of its largest element
    assert _err(pa.float(), a.double()) < 3e-7 and _err(pb.float(), b.double()) < 3e-7
    assert float(pa.amax) == float(a.abs().max())
    sa = float(pa.scale())
    assert 2 ** 13 <= float(a.abs().max()) * sa < 2 ** 14 and sa == 2.0 ** round(math.log2(sa))
    out = P.gemm(pa, pb, m, n, k, a_mn_major=a_mn, b_mn_major=b_mn, bias=bias)
    A = a.double().t() if a_mn else a.double()
    B = b.double() if b_mn else b.double().t()
    ref = A @ B + bias.double()
    assert _err(out, ref) < 6e-7, _err(out, ref)
    # split-K (weight-gradient style) applies the scales in the reduction kernel
    if a_mn and b_mn and k >= 256:
        out2 = P.gemm(pa, pb, m, n, k, a_mn_major=True, b_mn_major=True, split_k=3)
        assert _err(out2, A @ B) < 6e-7


def test_fp16x2_ones_column_yields_the_bias_gradient(LA):
    """gemm_dw over FP16X2 planes: the ones column stores the scale itself, so the extra output column is sum_r g[r, :]."""
    from recsys_benchmark_b200 import planes as P

    torch.manual_seed(1)
    g = torch.randn(4096, 400, device=DEV) * 1e-5
    x = torch.randn(4096, 624, device=DEV) * 0.03
    gp, xp = P.split(g, fmt=P.FP16X2), P.split(x, ones_col=True, fmt=P.FP16X2)
    assert xp.max_exp == P.ONES_SCALE_EXP and float(xp.scale()) <= 2.0 ** 14
    dw, db = P.gemm_dw(gp, xp, True)
    assert _err(dw, g.double().t() @ x.double()) < 6e-7
    assert _err(db, g.double().sum(0)) < 6e-7


@pytest.mark.parametrize("m,n,k", [(4096, 400, 624), (1000, 64, 352), (130, 48, 64), (65536, 400, 400), (32, 16, 32)])
@pytest.mark.parametrize("fmt", ["bf16x3", "fp16x2"])
def test_gemm_epilogue_batchnorm_statistics_match_the_separate_pass(LA, m, n, k, fmt):
    """rsb_gemm_epilogue.bn_partials: the GEMM epilogue reduces per-32-row-group shifted column sums of its output from
    the accumulator registers and rsb_bn_finalize_partials combines them (Chan) - same mean / rstd / affine / running
    buffers as the pass over z (rsb_bn_train_fwd_stats) and as fp64, incl. a ragged last row group; z itself is
    unchanged bit for bit."""
    from recsys_benchmark_b200 import planes as P

    F = P.FP16X2 if fmt == "fp16x2" else P.BF16X3
    torch.manual_seed(m + n)
    x = torch.randn(m, k, device=DEV) * 0.3 + 0.1
    w = torch.randn(n, k, device=DEV) * 0.05
    b = torch.randn(n, device=DEV) * 3.0                       # column means far from zero: |mean| >> std
    gamma, beta = torch.rand(n, device=DEV) + 0.5, torch.randn(n, device=DEV) * 0.2
    xp, wp = P.split(x, fmt=F), P.split(w, fmt=F)
    lib = LA.L.load()
    if lib.rsb_gemm_bn_partials_bytes(m, n, k, F) <= 0:
        pytest.skip("this shape does not take the TMA-store epilogue")
    rm1, rv1 = torch.zeros(n, device=DEV), torch.ones(n, device=DEV)
    rm2, rv2 = rm1.clone(), rv1.clone()
    a1 = torch.zeros(1, device=DEV)
    a2 = torch.zeros(1, device=DEV)
    z1, st1, af1 = P.gemm_bn_stats(xp, wp, b, gamma, beta, 1e-5, 0.1, rm1, rv1, act_amax=a1, bound_mul=2.0)
    z2 = P.gemm(xp, wp, m, n, k, bias=b, split_k=1)
    st2, af2 = P.bn_train_stats(z2, gamma, beta, 1e-5, 0.1, rm2, rv2, act_amax=a2, bound_mul=2.0)
    assert torch.equal(z1, z2)
    z64 = z2.double()
    mean64, var64 = z64.mean(0), z64.var(0, unbiased=False)
    rstd64 = 1.0 / torch.sqrt(var64 + 1e-5)
    for st in (st1, st2):
        assert _err(st[:n], mean64) < 2e-7
        assert float(((st[n:].double() - rstd64).abs() / rstd64).max()) < 3e-6
    assert float(((af1 - af2).abs() / (af2.abs() + 1e-3)).max()) < 1e-5
    assert _err(rm1, rm2.double()) < 1e-6 and _err(rv1, rv2.double()) < 1e-5
    assert float(a1) == float(a2) > 0


def test_gemm_alpha_beta_c(LA):
    a, b = torch.randn(256, 64, device=DEV), torch.randn(64, 128, device=DEV)
    c = torch.randn(256, 128, device=DEV)
    out = LA.gemm(a, b, alpha=0.5, beta=2.0, c=c)
    ref = 0.5 * (a.double() @ b.double()) + 2.0 * c.double()
    assert _err(out, ref) < 2e-6


def test_gemm_split_k_and_strided_batches(LA):
    a, b = torch.randn(65536, 400, device=DEV), torch.randn(65536, 624, device=DEV)
    out = LA.gemm(a, b, trans_a=True, split_k=32)              # a^T b : weight-gradient shape
    ref = a.double().t() @ b.double()
    assert _err(out, ref) < 2e-6
    h, c = torch.randn(2048, 4 * 64, device=DEV), torch.randn(4, 64, 64, device=DEV)
    out = LA.expert_matmul(h, c)
    ref = torch.einsum("ber,ers->bes", h.double().view(2048, 4, 64), c.double()).reshape(2048, 256)
    assert _err(out, ref) < 2e-6


def test_unsupported_shapes_fail_loudly_in_raw_call_and_dispatch_in_linear(LA):
    a, b = torch.randn(64, 30, device=DEV), torch.randn(30, 62, device=DEV)   # N = 62: fp32 results are stored 128 bits at a time
    with pytest.raises(RuntimeError):
        LA.gemm(a, b)
    b = torch.randn(30, 64, device=DEV)                                          # odd K is fine: operands become planes
    assert _err(LA.gemm(a, b), a.double() @ b.double()) < 2e-6
    w = torch.randn(1, 400, device=DEV)
    x = torch.randn(64, 400, device=DEV)
    torch.testing.assert_close(LA.linear(x, w, None), x @ w.t())   # N = 1 -> library GEMM


def test_linear_matmul_expert_autograd_match_torch(LA):
    torch.manual_seed(0)
    x = torch.randn(4096, 624, device=DEV, requires_grad=True)
    w = (torch.randn(400, 624, device=DEV) * 0.05).requires_grad_(True)
    b = torch.randn(400, device=DEV, requires_grad=True)
    gy = torch.randn(4096, 400, device=DEV)
    y = LA.linear(x, w, b)
    g1 = torch.autograd.grad(y, [x, w, b], gy)
    y2 = torch.nn.functional.linear(x.double(), w.double(), b.double())
    g2 = torch.autograd.grad(y2, [x, w, b], gy.double())
    assert _err(y, y2) < 2e-6
    for a_, b_ in zip(g1, g2):
        assert _err(a_, b_) < 2e-6
    a = torch.randn(2048, 256, device=DEV, requires_grad=True)
    m = torch.randn(256, 352, device=DEV, requires_grad=True)
    g = torch.randn(2048, 352, device=DEV)
    g1 = torch.autograd.grad(LA.matmul(a, m), [a, m], g)
    g2 = torch.autograd.grad(a.double() @ m.double(), [a, m], g.double())
    for a_, b_ in zip(g1, g2):
        assert _err(a_, b_) < 2e-6
    h = torch.randn(2048, 256, device=DEV, requires_grad=True)
    c = torch.randn(4, 64, 64, device=DEV, requires_grad=True)
    go = torch.randn(2048, 256, device=DEV)
    g1 = torch.autograd.grad(LA.expert_matmul(h, c), [h, c], go)
    ref = torch.einsum("ber,ers->bes", h.double().view(2048, 4, 64), c.double()).reshape(2048, 256)
    g2 = torch.autograd.grad(ref, [h, c], go.double())
    for a_, b_ in zip(g1, g2):
        assert _err(a_, b_) < 2e-6


def test_dcn_mix_head_avazu_shape_matches_reference_formulation():
    """Avazu-shaped cross head (Dm=352, E=4, r=64, L=3, B=2048) on the tensor-core kernel vs the
    reference's own formulation (src/models/layer_dcn.py:8-24,90-115) evaluated in fp64."""
    import recsys_benchmark_b200 as R

    torch.manual_seed(1)
    head = R.DCN_MixHead(4, 3, 64, 352).to(DEV)
    with torch.no_grad():
        for bia in head.biases:
            bia.normal_(0, 0.1)
    x0 = (torch.randn(2048, 352, device=DEV) * 0.3).requires_grad_(True)
    out = head(x0)
    gout = torch.randn_like(out)
    params = [x0] + list(head.parameters())
    g1 = torch.autograd.grad(out, params, gout)

    def ref_forward(x0d):
        xl = x0d
        x0u = x0d.unsqueeze(1)
        for l in range(3):
            V, C, U, bl = (head.V[l].double(), head.C[l].double(), head.U[l].double(), head.biases[l].double())
            E = torch.tanh(xl @ V).permute(1, 0, 2)
            E = torch.tanh(torch.einsum("ber,ers->bes", E, C))
            E = torch.einsum("ber,erd->bed", E, U)
            E = x0u * (E + bl)
            gts = (xl @ head.gates.double()).squeeze(2).permute(1, 0)
            xl = torch.einsum("be,bed->bd", gts, E) + xl
        return xl

    ref = ref_forward(x0.double())
    g2 = torch.autograd.grad(ref, params, gout.double())
    assert _err(out, ref) < 5e-6
    for a_, b_, p in zip(g1, g2, params):
        assert _err(a_, b_) < 2e-5, tuple(p.shape)


def test_fused_mlp_node_equals_the_unfused_composition(LA, monkeypatch):
    """[Linear -> ReLU -> Dropout] x 3 -> Linear(400, 1) as one autograd node (GEMM epilogues write the next GEMM's
    planes, the bias gradient rides the weight-gradient GEMM as a ones column) against the layer-by-layer path: the
    same Philox stream => the same masks, so the forward must agree bit for bit and the gradients to fp32 rounding;
    and against plain torch given those masks."""
    torch.manual_seed(0)
    seq = torch.nn.Sequential(torch.nn.Linear(624, 400), torch.nn.ReLU(), torch.nn.Dropout(0.5),
                              torch.nn.Linear(400, 400), torch.nn.ReLU(), torch.nn.Dropout(0.5),
                              torch.nn.Linear(400, 400), torch.nn.ReLU(), torch.nn.Dropout(0.5),
                              torch.nn.Linear(400, 1)).to(DEV).train()
    x = torch.randn(4096, 624, device=DEV, requires_grad=True)
    gout = torch.randn(4096, 1, device=DEV)
    params = [x] + list(seq.parameters())

    def run(fused):
        if not fused:
            monkeypatch.setattr(LA, "_mlp_relu_dropout_pattern", lambda mods, xx: None)
        monkeypatch.setattr(LA, "_DROPOUT_CALLS", 0)
        out = LA.run_sequential(seq, x)
        return out.detach(), torch.autograd.grad(out, params, gout)

    o1, g1 = run(True)
    o2, g2 = run(False)
    assert torch.equal(o1, o2)
    for a_, b_, p in zip(g1, g2, params):
        assert a_.shape == p.shape
        assert _err(a_, b_.double()) < 2e-6, tuple(p.shape)


def test_fused_epilogues_match_unfused_kernels(LA):
    """rsb_gemm_planes epilogue modes against GEMM + stand-alone passes: RELU_DROPOUT_PLANES (planes + mask) and
    MASK_PLANES / MASK_F32; rank-1 planes; the ones column of a weight-gradient GEMM is the bias gradient."""
    from recsys_benchmark_b200 import planes as P

    torch.manual_seed(1)
    m, k, n = 1000, 176, 400
    x = torch.randn(m, k, device=DEV)
    w = torch.randn(n, k, device=DEV) * 0.1
    b = torch.randn(n, device=DEV)
    xp, wp = P.split(x, ones_col=True), P.split(w)
    yp, mask = P.linear_relu_dropout(xp, wp, b, 0.3, seed=12345, offset=7 << 32)
    z = P.gemm(xp, wp, m, n, k, bias=b, split_k=1)
    lib = LA.L.load()
    y = torch.empty_like(z)
    mask2 = torch.empty(m, n, dtype=torch.uint8, device=DEV)
    LA.L.check(lib.rsb_relu_dropout_fwd(z.data_ptr(), z.numel(), 0.3, 12345, 7 << 32, None, y.data_ptr(), mask2.data_ptr(),
                                        LA.L.stream_ptr(DEV)))
    assert torch.equal(mask, mask2)
    yp2, mask3 = P.relu_dropout_planes(z, 0.3, seed=12345, offset=7 << 32)     # the one-pass un-fused form
    assert torch.equal(mask3, mask) and torch.equal(yp2.data, yp.data)
    assert torch.equal(yp.float(), P.split(y).float()) and abs(float(mask.float().mean()) - 0.35) < 0.02
    assert torch.equal(yp.data[0, :, n].float(), torch.ones(m, device=DEV)) and float(yp.data[1:, :, n:].abs().sum()) == 0
    # dX with the mask of the previous layer
    g = torch.randn(m, n, device=DEV)
    gp = P.split(g)
    mprev = (torch.rand(m, k, device=DEV) > 0.4).to(torch.uint8)
    ref = (g.double() @ w.double()) * mprev.double() / 0.7
    wtp = P.split(w, transpose=True)
    assert _err(P.dx_masked(gp, wtp, mprev, 0.3, to_planes=False), ref) < 2e-6
    assert _err(P.dx_masked(gp, wtp, mprev, 0.3).float(), ref) < 2e-6
    # weight gradient + bias gradient from the ones column
    dw, db = P.gemm_dw(gp, xp, True)
    assert _err(dw, g.double().t() @ x.double()) < 2e-6 and _err(db, g.double().sum(0)) < 2e-6
    # rank-1 upstream gradient
    gr, wc = torch.randn(m, device=DEV), torch.randn(n, device=DEV)
    r1 = P.rank1_mask_planes(gr, wc, mask, 0.3).float()
    assert _err(r1, gr.double()[:, None] * wc.double()[None, :] * mask.double() / 0.7) < 1e-6


def test_gemm_speed_report(LA, capsys):
    """Not an assertion on speed: prints TFLOP/s of the tensor-core kernel vs cuBLAS fp32 for the MLP shapes."""
    res = []
    for (m, n, k, ta, tb, sk) in [(65536, 400, 624, False, True, 1), (65536, 400, 400, False, True, 1),
                                  (65536, 624, 400, False, False, 1), (400, 624, 65536, True, False, 0),
                                  (65536, 256, 352, False, False, 1), (65536, 352, 256, False, False, 1)]:
        a = torch.randn((k, m) if ta else (m, k), device=DEV)
        b = torch.randn((n, k) if tb else (k, n), device=DEV)
        at, bt = (a.t() if ta else a), (b.t() if tb else b)
        for fn, name in [(lambda: LA.gemm(a, b, trans_a=ta, trans_b=tb, split_k=sk), "rsb"), (lambda: at @ bt, "cublas")]:
            for _ in range(3):
                fn()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            res.append((m, n, k, name, ms, 2 * m * n * k / ms / 1e9))
    with capsys.disabled():
        for r in res:
            print("GEMM %6d x %4d x %6d %-6s %8.3f ms %7.1f TFLOP/s" % r)


def test_colsum_matches_fp64(LA):
    x = torch.randn(65536, 400, device=DEV)
    ref = x.double().sum(0)
    assert _err(LA.colsum(x), ref) < 1e-6
    assert torch.equal(LA.colsum(x), LA.colsum(x))          # deterministic
    y = torch.randn(1000, 2400, device=DEV)                 # more than one column pass
    assert _err(LA.colsum(y), y.double().sum(0)) < 1e-6


def test_relu_dropout_statistics_and_backward(LA):
    torch.manual_seed(3)
    x = torch.randn(8192, 400, device=DEV, requires_grad=True)
    y = LA.relu_dropout(x, 0.5, True)
    pos = x.detach() > 0
    assert float(y.detach()[~pos].abs().sum()) == 0.0
    kept = y.detach() != 0
    assert torch.equal(y.detach()[kept], (2.0 * x.detach())[kept])
    frac = float(kept.sum()) / float(pos.sum())
    assert abs(frac - 0.5) < 0.01, frac
    g = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, g)
    assert torch.equal(gx, torch.where(kept, 2.0 * g, torch.zeros_like(g)))
    y2 = LA.relu_dropout(x, 0.5, True)
    assert not torch.equal(y2, y)                            # a new stream every call
    torch.testing.assert_close(LA.relu_dropout(x, 0.5, False), torch.relu(x))   # eval mode = relu


def test_fused_linear_relu_dropout_block_matches_composition(LA):
    torch.manual_seed(4)
    seq = torch.nn.Sequential(torch.nn.Linear(624, 400), torch.nn.ReLU(), torch.nn.Dropout(0.5),
                              torch.nn.Linear(400, 400), torch.nn.BatchNorm1d(400), torch.nn.ReLU(),
                              torch.nn.Dropout(0.2), torch.nn.Linear(400, 1)).to(DEV).train()
    x = torch.randn(4096, 624, device=DEV, requires_grad=True)
    out = LA.run_sequential(seq, x)
    assert tuple(out.shape) == (4096, 1)
    go = torch.randn_like(out)
    params = [x] + list(seq.parameters())
    g1 = torch.autograd.grad(out, params, go)
    # first block in isolation against an fp64 composition that reuses the mask the kernel drew
    lin = seq[0]
    y = LA._LinearReluDropout.apply(x, lin.weight, lin.bias, 0.5)
    keep = (y.detach() != 0).double()
    z = torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double())
    ref = z * keep * 2.0
    assert _err(y, ref) < 2e-6
    gy = torch.randn_like(y)
    ga = torch.autograd.grad(y, [x, lin.weight, lin.bias], gy)
    gb = torch.autograd.grad(ref, [x, lin.weight, lin.bias], gy.double())
    for a_, b_ in zip(ga, gb):
        assert _err(a_, b_) < 2e-6
    assert all(torch.isfinite(g).all() for g in g1)
    seq.eval()
    torch.testing.assert_close(LA.run_sequential(seq, x), seq(x), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("with_bn", [False, True])
def test_head_block_matches_unfused_path_with_the_same_masks(LA, with_bn, monkeypatch):
    """[Linear ->] ReLU -> Dropout -> Linear(hidden, 1): the fused head (one-output Linear folded into the glue
    passes) against the unfused composition, with the Philox stream rewound so both draw identical masks."""
    torch.manual_seed(5)
    layers = [torch.nn.Linear(624, 400)]
    if with_bn:
        layers.append(torch.nn.BatchNorm1d(400))
    layers += [torch.nn.ReLU(), torch.nn.Dropout(0.5), torch.nn.Linear(400, 1)]
    seq = torch.nn.Sequential(*layers).to(DEV).train()
    x = torch.randn(8192, 624, device=DEV, requires_grad=True)
    params = [x] + [p for p in seq.parameters()]
    go = torch.randn(8192, 1, device=DEV)

    # (the no-BatchNorm sequence would be taken whole by the fused MLP node: this test is about the head block)
    monkeypatch.setattr(LA, "_mlp_relu_dropout_pattern", lambda mods, xx: None)
    monkeypatch.setattr(LA, "_mlp_batchnorm_pattern", lambda mods, xx: None)
    calls = LA._DROPOUT_CALLS
    out_f = LA.run_sequential(seq, x)
    assert out_f.grad_fn.name().startswith("_HeadBlock")
    g_f = torch.autograd.grad(out_f, params, go)

    LA._DROPOUT_CALLS = calls                          # same (seed, offset) -> same masks
    monkeypatch.setattr(LA, "_is_head", lambda mod, width: False)
    out_u = LA.run_sequential(seq, x)
    assert not out_u.grad_fn.name().startswith("_HeadBlock")
    g_u = torch.autograd.grad(out_u, params, go)

    assert _err(out_f, out_u.double()) < 2e-6
    names = ["x"] + [n for n, _ in seq.named_parameters()]
    for n, a_, b_ in zip(names, g_f, g_u):
        if with_bn and n == "0.bias":
            continue                                   # zero by construction in front of BatchNorm: fp32 noise
        assert _err(a_, b_.double()) < 5e-6, n
    # fp64 check of the one-output Linear itself on the activations the kernel produced
    y, mask, out = LA._relu_dropout_dot_fwd(x.detach()[:, :400].contiguous(), 0.5, seq[-1].weight.reshape(-1),
                                            seq[-1].bias)
    ref = y.double() @ seq[-1].weight.double().t() + seq[-1].bias.double()
    assert _err(out, ref[:, 0]) < 2e-6
    keep = mask.float().mean().item()
    assert abs(keep - 0.25) < 0.01                     # P(x > 0) * (1 - p)


@pytest.mark.parametrize("fmt", ["bf16x3", "fp16x2"])
@pytest.mark.parametrize("p", [0.0, 0.5])
def test_fused_batchnorm_mlp_node_matches_torch(LA, monkeypatch, p, fmt):
    """[Linear -> BatchNorm1d -> ReLU -> Dropout] x 3 -> Linear(400, 1) as one autograd node (own batch statistics,
    BatchNorm + ReLU + dropout -> planes passes, BatchNorm backward -> planes) against torch's own modules in fp64
    driven with the masks the fused node drew; running statistics follow torch's update rule."""
    import copy

    from recsys_benchmark_b200 import planes as P

    monkeypatch.setattr(LA, "MLP_PLANES_FORMAT", P.FP16X2 if fmt == "fp16x2" else P.BF16X3)
    torch.manual_seed(0)
    mods = []
    width = 624
    for _ in range(3):
        mods += [torch.nn.Linear(width, 400), torch.nn.BatchNorm1d(400), torch.nn.ReLU(), torch.nn.Dropout(p)]
        width = 400
    mods.append(torch.nn.Linear(400, 1))
    seq = torch.nn.Sequential(*mods).to(DEV).train()
    with torch.no_grad():
        for m in seq:
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    ref = copy.deepcopy(seq).double()
    x = torch.randn(4096, 624, device=DEV, requires_grad=True)
    gout = torch.randn(4096, 1, device=DEV)
    out = LA.run_sequential(seq, x)
    assert type(out.grad_fn).__name__ == "_MlpBatchNormBackward"
    masks = out.grad_fn.masks                      # uint8 keep-and-positive masks, one per layer
    params = [x] + list(seq.parameters())
    g1 = torch.autograd.grad(out, params, gout)
    # fp64 reference with the SAME ReLU + dropout decisions (the keep-and-positive masks of the fused node), so that a
    # pre-activation within fp32 rounding of zero cannot land on different sides: y = bn(z) * mask / (1 - p)
    h = x.double()
    keep_iter = iter(masks)
    for m in ref:
        if isinstance(m, torch.nn.ReLU):
            continue
        if isinstance(m, torch.nn.Dropout):
            h = h * next(keep_iter).double() / (1.0 - p)
        else:
            h = m(h)
    g2 = torch.autograd.grad(h, [x] + list(ref.parameters()), gout.double(), allow_unused=True)
    assert _err(out, h) < 5e-6
    names = ["x"] + [n for n, _ in seq.named_parameters()]
    for n, a_, b_ in zip(names, g1, g2):
        if n.endswith(".bias") and n.split(".")[0] in ("0", "4", "8"):
            assert float(a_.abs().max()) < 1e-4      # Linear bias in front of BatchNorm: zero by construction
            continue
        assert _err(a_, b_) < 2e-5, n
    for ms, mr in zip(seq, ref):
        if isinstance(ms, torch.nn.BatchNorm1d):
            assert _err(ms.running_mean, mr.running_mean) < 1e-5 and _err(ms.running_var, mr.running_var) < 1e-5
            assert int(ms.num_batches_tracked) == 1
