"""fp32-accurate tensor-core GEMMs (rsb_gemm_planes: hand-written TMA + tcgen05 + TMEM kernel on operands held as
three bf16 planes, see planes.py / csrc/gemm/planes_gemm.cu) behind `matmul` / `linear` autograd functions.

`gemm()` is the raw fp32 call (row-major operands with optional "stored transposed" flags, fused alpha / beta*C /
per-column bias).  Results narrower than 4 columns or not a multiple of 4 wide (the final Linear(400->1), the toy
sizes of the unit tests) go to the library GEMM (torch.matmul -> cuBLAS fp32); that is a shape dispatch between two
fp32 GEMMs, not a CPU fallback."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import functional as RF
from . import planes as P

MIN_FLOPS = 1 << 22   # below this a library call is launch-latency-equivalent


def _aligned(*tensors) -> bool:
    return all(t is None or t.data_ptr() % 16 == 0 for t in tensors)


def gemm_supported(m: int, n: int, k: int, lda: int, ldb: int, ldd: int, trans_a: bool, trans_b: bool) -> bool:
    """Shapes the tensor-core kernel takes: fp32 results are stored 128 bits at a time (N and the output leading
    dimension multiples of 4); operands are re-laid out as bf16 planes, so their shapes are free."""
    if trans_a and trans_b:
        return False
    return min(m, n, k) > 0 and n % 4 == 0 and ldd % 4 == 0


def weight_planes(w: torch.Tensor, transpose: bool = False, fmt: int = P.BF16X3) -> P.Planes:
    """Planes of a parameter (or of its transpose), rebuilt only when the parameter has changed (in-place updates bump
    `_version`): one split per optimizer step.  The forward GEMM reads W [out, in] as the K-major operand; the dX GEMM
    reads the planes of W^T [in, out], K-major as well (the MN-major view of the same planes would also do, but its
    shared-memory tiles are padded to 64 columns, which costs the room the TMA-store epilogue needs).
    FP16X2: W and W^T share one max |W| scalar (one rsb_absmax per version)."""
    attr = ("_rsb_planes_t" if transpose else "_rsb_planes") + ("_h" if fmt == P.FP16X2 else "")
    cached = getattr(w, attr, None)
    key = (w._version, w.data_ptr(), tuple(w.shape))
    if cached is not None and cached[0] == key:
        return cached[1]
    w2 = w.detach().reshape(-1, w.shape[-1])
    amax = None
    if fmt == P.FP16X2:
        other = getattr(w, ("_rsb_planes_h" if transpose else "_rsb_planes_t_h"), None)
        amax = other[1].amax if (other is not None and other[0] == key) else P.absmax(w2)
    pl = P.split(w2, transpose=transpose, fmt=fmt, amax=amax)
    setattr(w, attr, (key, pl))
    return pl


# Operand format of the fused training-mode dense tails (_MlpBatchNorm): FP16X2 halves the MMAs and the plane bytes;
# set to P.BF16X3 for the exact-operand path.
MLP_PLANES_FORMAT = P.FP16X2
# BatchNorm batch statistics reduced in the GEMM epilogue (rsb_gemm_epilogue.bn_partials) instead of by a pass over z.
# Correct and tested, but measured no faster on the headline step: the statistics lengthen the GEMM's exposed epilogue
# by as much (+0.024 ms per GEMM, + 0.023 ms to combine 2048 row groups) as the separate pass costs (0.050 ms).  Off.
BN_STATS_IN_GEMM = False


def gemm(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False,
         bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, alpha: float = 1.0,
         beta: float = 0.0, c: Optional[torch.Tensor] = None, split_k: int = 1,
         a_planes: Optional[P.Planes] = None, b_planes: Optional[P.Planes] = None) -> torch.Tensor:
    """D = alpha * op(a) @ op(b) + beta * c + bias for 2-D fp32 CUDA tensors.

    trans_a: `a` is stored [K, M]; trans_b: `b` is stored [N, K].  The operands are split into bf16 planes (or taken
    from `a_planes` / `b_planes` when the caller already holds them) and multiplied by rsb_gemm_planes; no operand is
    ever transposed in memory - the kernel reads either majorness.  split_k != 1 lets the kernel split the reduction
    over the SMs (weight-gradient GEMMs whose K is the batch); partials are summed in fixed order."""
    dev = L.require_cuda(a, b, bias, c)
    assert a.dim() == 2 and b.dim() == 2 and a.dtype == torch.float32 and b.dtype == torch.float32
    k, m = (a.shape if trans_a else a.shape[::-1])
    kb, n = (b.shape[::-1] if trans_b else b.shape)
    assert k == kb, f"inner dimensions differ: {k} vs {kb}"
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=dev)
    if not gemm_supported(m, n, k, 0, 0, out.stride(0), trans_a, trans_b) or out.stride(1) != 1 or \
            not _aligned(bias, c, out):
        raise RuntimeError("rsb gemm: shape not supported by the tensor-core kernel (N and ldd must be multiples of 4)")
    pa = a_planes if a_planes is not None else P.split(a)
    pb = b_planes if b_planes is not None else P.split(b)
    return P.gemm(pa, pb, m, n, k, a_mn_major=trans_a, b_mn_major=not trans_b, bias=bias, out=out, alpha=alpha,
                  beta=beta, c=c, split_k=1 if split_k == 1 else 0)


def _fwd_gemm(xp: P.Planes, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """x W^T + b for W [out, in]: both operands K-major (the weight planes in the activation's format)."""
    n, k = weight.shape
    return P.gemm(xp, weight_planes(weight, fmt=xp.fmt), xp.rows, n, k, bias=bias, split_k=1)


def _dx_gemm(gp: P.Planes, weight: torch.Tensor) -> torch.Tensor:
    """g W for W [out, in]: B = the planes of W^T [in, out], K-major."""
    n_out, n_in = weight.shape
    return P.gemm(gp, weight_planes(weight, transpose=True, fmt=gp.fmt), gp.rows, n_in, n_out, split_k=1)


def _dw_gemm(gp: P.Planes, xp: P.Planes) -> torch.Tensor:
    """g^T x: the reduction runs over the stored rows (the batch) of both operands; split over the SMs."""
    return P.gemm(gp, xp, gp.cols, xp.cols, gp.rows, a_mn_major=True, b_mn_major=True, split_k=0)


class _ExpertMatMul(torch.autograd.Function):
    """Block-diagonal product of the DCN-Mix experts (src/models/layer_dcn.py:22):
    out[b, e, :] = h[b, e, :] @ C[e]   for h [B, E*r] (expert-major columns), C [E, r, r2].
    Three batched tensor-core GEMMs (batch = experts, addressed as column / row blocks inside the stored matrices):
    forward, dh, and the split-K weight gradient; h and g are split into planes once each."""

    @staticmethod
    def forward(ctx, h, c):
        e, r, r2 = c.shape
        bsz = h.shape[0]
        hp = P.split(h)
        cp = weight_planes(c)                                   # stored [E*r, r2]
        out = torch.empty(bsz, e * r2, dtype=torch.float32, device=h.device)
        # A = h[:, l*r:(l+1)*r] (K-major, column step r); B = C[l] stored [K = r, N = r2] (MN-major, row step r)
        P.gemm(hp, cp, bsz, r2, r, a_mn_major=False, b_mn_major=True, out=out, batch=e, a_steps=(0, r),
               b_steps=(r, 0), d_batch_stride=r2, split_k=1)
        ctx.hp = hp
        ctx.save_for_backward(c)
        ctx.shape = (bsz, e, r, r2)
        return out

    @staticmethod
    def backward(ctx, g):
        (c,) = ctx.saved_tensors
        bsz, e, r, r2 = ctx.shape
        hp, ctx.hp = ctx.hp, None
        gp = P.split(g)
        gh = gc = None
        if ctx.needs_input_grad[0]:
            gh = torch.empty(bsz, e * r, dtype=torch.float32, device=g.device)
            # gh[:, l, :] = g[:, l, :] @ C[l]^T : B = C[l] read as stored [N = r, K = r2] (K-major, row step r)
            P.gemm(gp, weight_planes(c), bsz, r, r2, a_mn_major=False, b_mn_major=False, out=gh, batch=e,
                   a_steps=(0, r2), b_steps=(r, 0), d_batch_stride=r, split_k=1)
        if ctx.needs_input_grad[1]:
            # C_grad[l] = h[:, l, :]^T @ g[:, l, :] : both operands MN-major (K = batch rows), column steps r / r2
            gc = torch.empty(e, r, r2, dtype=torch.float32, device=g.device)
            P.gemm(hp, gp, r, r2, bsz, a_mn_major=True, b_mn_major=True, out=gc.view(e * r, r2), batch=e,
                   a_steps=(0, r), b_steps=(0, r2), d_batch_stride=r * r2, split_k=0)
        return gh, gc


def expert_matmul(h: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """h [B, E*r] x C [E, r, r2] -> [B, E*r2] (block diagonal)."""
    e, r, r2 = c.shape
    bsz = h.shape[0]
    if r % 32 == 0 and r2 % 4 == 0 and 2 * bsz * e * r * r2 >= MIN_FLOPS and h.is_cuda and h.dtype == torch.float32:
        return _ExpertMatMul.apply(h, c)
    return torch.bmm(h.view(bsz, e, r).transpose(0, 1), c).transpose(0, 1).reshape(bsz, e * r2)


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a 2-D fp32 CUDA tensor (deterministic two-stage kernel); torch.sum for odd shapes."""
    if x.dim() == 2 and x.is_cuda and x.shape[1] % 4 == 0 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and \
            x.data_ptr() % 16 == 0 and x.shape[0] >= 256:
        lib = L.load()
        m, n = x.shape
        out = torch.empty(n, dtype=torch.float32, device=x.device)
        ws = RF._ws(lib.rsb_colsum_workspace_bytes(m, n), x.device)
        RF._call("colsum", lib.rsb_colsum, L.ptr(x), m, n, x.stride(0), L.ptr(out), L.ptr(ws), ws.numel(),
                 L.stream_ptr(x.device), nbytes=m * n * 4)
        return out
    return x.sum(0)


_DROPOUT_CALLS = 0
_DROPOUT_DEV_COUNTER = None   # int64 device tensor: Philox stream position for CUDA-graph replays


def use_device_dropout_counter(device) -> torch.Tensor:
    """CUDA-graph mode: the dropout stream position is read from device memory, so that every replay
    of a captured step draws fresh masks.  Call `advance_dropout_counter()` once per step INSIDE the
    captured region."""
    global _DROPOUT_DEV_COUNTER
    _DROPOUT_DEV_COUNTER = torch.zeros(1, dtype=torch.int64, device=device)
    return _DROPOUT_DEV_COUNTER


def advance_dropout_counter():
    if _DROPOUT_DEV_COUNTER is not None:
        _DROPOUT_DEV_COUNTER.add_(1 << 40)


def _dropout_stream(numel: int):
    """(seed, offset) of the next Philox stream: seeded by torch's global seed, advanced per call."""
    global _DROPOUT_CALLS
    _DROPOUT_CALLS += 1
    seed = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + _DROPOUT_CALLS * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF
    return seed, (_DROPOUT_CALLS * 0x100000000) & 0xFFFFFFFFFFFFFFFF


def _relu_dropout_fwd(x: torch.Tensor, p: float):
    lib = L.load()
    y = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    seed, off = _dropout_stream(x.numel())
    RF._call("relu_dropout_fwd", lib.rsb_relu_dropout_fwd, L.ptr(x), x.numel(), float(p), seed, off,
             L.ptr(_DROPOUT_DEV_COUNTER), L.ptr(y), L.ptr(mask), L.stream_ptr(x.device), nbytes=x.numel() * 9)
    return y, mask


def _relu_dropout_bwd(g: torch.Tensor, mask: torch.Tensor, p: float, want_colsum: bool):
    lib = L.load()
    m, n = g.shape
    g = g.contiguous()
    gx = torch.empty_like(g)
    cs = torch.empty(n, dtype=torch.float32, device=g.device) if want_colsum else None
    ws = RF._ws(lib.rsb_colsum_workspace_bytes(m, n), g.device) if want_colsum else None
    RF._call("relu_dropout_bwd", lib.rsb_relu_dropout_bwd, L.ptr(g), L.ptr(mask), m, n, float(p), L.ptr(gx), L.ptr(cs),
             L.ptr(ws), ws.numel() if ws is not None else 0, L.stream_ptr(g.device), nbytes=m * n * 9)
    return gx, cs


def _relu_dropout_dot_fwd(x: torch.Tensor, p: float, w: torch.Tensor, bias: Optional[torch.Tensor],
                          affine: Optional[torch.Tensor] = None):
    """(y, mask, out) with out[r] = dropout(relu(x))[r,:] . w + bias: the one-output Linear folded into the pass.
    affine [2N]: x is first normalised as x * affine[:N] + affine[N:] (BatchNorm folded in as well)."""
    lib = L.load()
    m, n = x.shape
    y = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    out = torch.empty(m, dtype=torch.float32, device=x.device)
    seed, off = _dropout_stream(x.numel())
    RF._call("relu_dropout_dot_fwd", lib.rsb_relu_dropout_dot_fwd, L.ptr(x), m, n, float(p), seed, off,
             L.ptr(_DROPOUT_DEV_COUNTER), L.ptr(w), L.ptr(bias), L.ptr(affine), L.ptr(y), L.ptr(mask), L.ptr(out),
             L.stream_ptr(x.device), nbytes=x.numel() * 9 + m * 4)
    return y, mask, out


def _relu_dropout_bwd_rank1(g_row: torch.Tensor, w_col: torch.Tensor, mask: torch.Tensor, p: float, want_colsum: bool):
    """relu_dropout_bwd of the upstream gradient g_row[r] * w_col[c] (never materialised)."""
    lib = L.load()
    m, n = mask.shape
    gx = torch.empty(m, n, dtype=torch.float32, device=mask.device)
    cs = torch.empty(n, dtype=torch.float32, device=mask.device) if want_colsum else None
    ws = RF._ws(lib.rsb_colsum_workspace_bytes(m, n), mask.device) if want_colsum else None
    RF._call("relu_dropout_bwd_rank1", lib.rsb_relu_dropout_bwd_rank1, L.ptr(g_row), L.ptr(w_col), L.ptr(mask), m, n,
             float(p), L.ptr(gx), L.ptr(cs), L.ptr(ws), ws.numel() if ws is not None else 0,
             L.stream_ptr(mask.device), nbytes=m * n * 5 + m * 4)
    return gx, cs


def _colsum_weighted(x: torch.Tensor, row_weight: torch.Tensor) -> torch.Tensor:
    """out[c] = sum_r row_weight[r] * x[r,c]."""
    lib = L.load()
    m, n = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    ws = RF._ws(lib.rsb_colsum_workspace_bytes(m, n), x.device)
    RF._call("colsum_weighted", lib.rsb_colsum_weighted, L.ptr(x), L.ptr(row_weight), m, n, x.stride(0), L.ptr(out),
             L.ptr(ws), ws.numel(), L.stream_ptr(x.device), nbytes=m * n * 4 + m * 4)
    return out


def _fusable(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous() and x.shape[1] % 4 == 0
            and x.shape[1] <= 2048 and x.data_ptr() % 16 == 0)


class _ReluDropout(torch.autograd.Function):
    """y = dropout_p(relu(x)) in one pass each way (torch: clamp + fused_dropout / masked_scale + threshold_backward)."""

    @staticmethod
    def forward(ctx, x, p):
        y, mask = _relu_dropout_fwd(x, p)
        ctx.save_for_backward(mask)
        ctx.p = p
        return y

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        gx, _ = _relu_dropout_bwd(g, mask, ctx.p, False)
        return gx, None


def relu_dropout(x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    if training and 0.0 < p < 1.0 and _fusable(x):
        return _ReluDropout.apply(x, p)
    return torch.nn.functional.dropout(torch.relu(x), p, training)


class _LinearReluDropout(torch.autograd.Function):
    """dropout_p(relu(x @ W^T + b)): tensor-core GEMM with the bias in its epilogue, one fused
    ReLU+dropout pass, and a backward whose single elementwise pass also produces the bias gradient.
    The planes of x are shared by the forward and the weight-gradient GEMM, those of gz by dX and dW."""

    @staticmethod
    def forward(ctx, x, weight, bias, p, side_dw=False):
        xp = P.split(x)
        z = _fwd_gemm(xp, weight, bias)
        y, mask = _relu_dropout_fwd(z, p)
        ctx.xp = xp
        ctx.save_for_backward(weight, mask)
        ctx.p = p
        ctx.has_bias = bias is not None
        ctx.side_dw = side_dw
        return y

    @staticmethod
    def backward(ctx, gy):
        weight, mask = ctx.saved_tensors
        xp, ctx.xp = ctx.xp, None
        gz, gb = _relu_dropout_bwd(gy, mask, ctx.p, ctx.has_bias and ctx.needs_input_grad[2])
        gp = P.split(gz)
        gx = _dx_gemm(gp, weight) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            if ctx.side_dw and gx is not None and _side_dw_safe(weight):
                gw = _dw_on_side_stream(gp, xp)
            else:
                gw = _dw_gemm(gp, xp)
        return gx, gw, gb, None, None


def _has_hooks(weight: torch.Tensor) -> bool:
    return bool(getattr(weight, "_backward_hooks", None)) or bool(getattr(weight, "_post_accumulate_grad_hooks", None))


def _side_dw_safe(weight: torch.Tensor) -> bool:
    """The side-stream weight gradient is handed to autograd while its GEMM is still running; the main stream only
    re-joins in the end-of-backward callback.  That is safe only if nothing touches the gradient before then:
    AccumulateGrad must STEAL it (`weight.grad is None`: no `grad += gw` on the main stream - gradient accumulation,
    zero_grad(set_to_none=False) or a second backward all make .grad non-None) and no tensor / post-accumulate hook
    may read it.  Otherwise the GEMM runs in line."""
    return weight.grad is None and not _has_hooks(weight) and not torch.cuda.is_current_stream_capturing()


def _dw_on_side_stream(gp: P.Planes, xp: P.Planes) -> torch.Tensor:
    """Weight gradient gz^T @ x of the FIRST dense layer on the side stream: everything queued after it in the
    backward pass is the embedding side (row gradients, segmented reduction - HBM/latency-bound kernels that
    leave the tensor pipe idle), so the two overlap.  The main stream re-joins at the end of the backward pass
    (autograd end-of-pass callback), before any optimizer can read the result."""
    dev = gp.data.device
    main = torch.cuda.current_stream(dev)
    side = RF.side_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        gw = _dw_gemm(gp, xp)
    # the callback's closure keeps the main-pool inputs alive until the main stream has re-joined (no
    # record_stream: it makes the caching allocator poll events and cudaMalloc when the host runs ahead)
    try:
        torch.autograd.Variable._execution_engine.queue_callback(lambda keep=(gp, xp): main.wait_stream(side))
    except Exception:  # noqa: BLE001 - not inside an autograd pass (or the hook is gone): join right away
        main.wait_stream(side)
    return gw


class _HeadBlock(torch.autograd.Function):
    """The tail of the MLPs, [Linear ->] ReLU -> Dropout -> Linear(hidden, 1) (src/models/deepfm.py:55-66,
    src/models/dcn.py:56-66), with the one-output Linear folded into the glue passes:
      fwd  z = x @ W^T + b (tensor-core GEMM, only when W is given), then ONE pass: y = dropout(relu(z)), mask,
           out = y . w_out + b_out;
      bwd  dW_out = sum_r g[r] y[r,:] (row-weighted column sum), db_out = sum g, then ONE pass forming
           g[r] * w_out[c] on the fly -> gz (+ its column sums = db), then the two GEMMs of the Linear.
    Returns out [M, 1].  With W = None the input is the pre-activation itself (BatchNorm sits in between)."""

    @staticmethod
    def forward(ctx, x, weight, bias, p, w_out, b_out):
        ctx.xp = None
        if weight is None:
            z = x
        else:
            ctx.xp = P.split(x)
            z = _fwd_gemm(ctx.xp, weight, bias)
        y, mask, out = _relu_dropout_dot_fwd(z, p, w_out.reshape(-1), b_out)
        ctx.save_for_backward(weight, mask, y, w_out)
        ctx.p = p
        ctx.has_bias = bias is not None
        ctx.has_bout = b_out is not None
        return out.unsqueeze(1)

    @staticmethod
    def backward(ctx, g_out):
        weight, mask, y, w_out = ctx.saved_tensors
        xp, ctx.xp = ctx.xp, None
        g = g_out.reshape(-1).contiguous()
        g_wout = _colsum_weighted(y, g).reshape(w_out.shape) if ctx.needs_input_grad[4] else None
        g_bout = g.sum().reshape(1) if ctx.has_bout and ctx.needs_input_grad[5] else None
        want_gb = weight is not None and ctx.has_bias and ctx.needs_input_grad[2]
        gz, gb = _relu_dropout_bwd_rank1(g, w_out.reshape(-1), mask, ctx.p, want_gb)
        if weight is None:
            return gz, None, None, None, g_wout, g_bout
        gp = P.split(gz)
        gx = _dx_gemm(gp, weight) if ctx.needs_input_grad[0] else None
        gw = _dw_gemm(gp, xp) if ctx.needs_input_grad[1] else None
        return gx, gw, gb, None, g_wout, g_bout


def _use_kernel(m, n, k, *tensors) -> bool:
    """The three GEMMs of a Linear produce [m,n], [m,k] and [n,k] fp32 results: their widths must be multiples of 4."""
    return (2 * m * n * k >= MIN_FLOPS and n % 4 == 0 and k % 4 == 0 and _aligned(*tensors)
            and all(t is None or (t.is_cuda and t.dtype == torch.float32) for t in tensors))


class _Linear(torch.autograd.Function):
    """y = x @ W^T + b with W [out, in] (nn.Linear layout); all three GEMMs on the tensor-core kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, side_dw=False):
        xp = P.split(x)
        ctx.xp = xp
        ctx.save_for_backward(weight)
        ctx.has_bias = bias is not None
        ctx.side_dw = side_dw
        return _fwd_gemm(xp, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        (weight,) = ctx.saved_tensors
        xp, ctx.xp = ctx.xp, None
        gy = gy.contiguous()
        gp = P.split(gy)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _dx_gemm(gp, weight)                               # [B,out] @ [out,in]
        if ctx.needs_input_grad[1]:
            if ctx.side_dw and gx is not None and _side_dw_safe(weight):
                gw = _dw_on_side_stream(gp, xp)
            else:
                gw = _dw_gemm(gp, xp)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = colsum(gy)
        return gx, gw, gb, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           side_dw: bool = False) -> torch.Tensor:
    """Drop-in for F.linear on 2-D inputs: tensor-core kernel when the shape allows, cuBLAS otherwise."""
    if x.dim() == 2 and _use_kernel(x.shape[0], weight.shape[0], weight.shape[1], x, weight, bias) and \
            x.stride(1) == 1 and x.stride(0) % 4 == 0:
        return _Linear.apply(x, weight, bias, side_dw)
    return torch.nn.functional.linear(x, weight, bias)


class _MatMul(torch.autograd.Function):
    """c = a @ b for 2-D row-major a [M,K], b [K,N]."""

    @staticmethod
    def forward(ctx, a, b):
        ap = P.split(a)
        ctx.ap = ap
        ctx.save_for_backward(b)
        m, k = a.shape
        return P.gemm(ap, weight_planes(b), m, b.shape[1], k, b_mn_major=True, split_k=1)

    @staticmethod
    def backward(ctx, gc):
        (b,) = ctx.saved_tensors
        ap, ctx.ap = ctx.ap, None
        gp = P.split(gc.contiguous())
        ga = gb = None
        if ctx.needs_input_grad[0]:
            # ga = gc @ b^T: b stored [K, N] read as the K-major "[N' = K, K' = N]" operand
            ga = P.gemm(gp, weight_planes(b), gp.rows, b.shape[0], b.shape[1], split_k=1)
        if ctx.needs_input_grad[1]:
            gb = P.gemm(ap, gp, ap.cols, gp.cols, ap.rows, a_mn_major=True, b_mn_major=True, split_k=0)   # a^T @ gc
        return ga, gb


def matmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1 and a.stride(0) % 4 == 0 and \
            b.stride(0) % 4 == 0 and _use_kernel(a.shape[0], b.shape[1], a.shape[1], a, b):
        return _MatMul.apply(a, b)
    return torch.matmul(a, b)


# The GEMM's epilogue runs after the tile's last K group, on the warps that hold the accumulators: it is not hidden
# behind the next tile's MMAs.  Writing planes + mask from there (each thread owns one row: 16-byte pieces of 32
# different rows per store instruction) measured 0.39 ms for 65536 x 400 x 400 against 0.20 ms for the plain fp32
# epilogue + 0.05 ms for the coalesced ReLU + dropout -> planes pass, so the forward keeps the two-launch form; the
# backward's masked dX epilogue (no mask write) is on par with the split form and stays fused.
FUSE_FORWARD_EPILOGUE = False


class _MlpReluDropout(torch.autograd.Function):
    """The whole dense tail  [Linear -> ReLU -> Dropout] x L  ->  Linear(hidden, 1)  (src/models/deepfm.py:55-66 with
    use_batchnorm off) as ONE autograd node whose activations live as bf16 planes between the tensor-core GEMMs:

      fwd  layer i < L : ONE launch - GEMM, bias, ReLU, dropout in the epilogue, output written as the next GEMM's planes
                         (+ the 1-byte mask); no fp32 activation is ever materialised
           layer L     : GEMM -> fp32, then the pass that applies ReLU + dropout AND the one-output Linear's dot product
      bwd  head        : dW_out, db_out, and g[r] * w_out[c] * mask as planes (one pass)
           layer i     : dX GEMM whose epilogue applies the previous layer's mask and writes planes (the operand of that
                         layer's dX and dW GEMMs); dW GEMM g^T [x | 1] over the planes both directions already hold,
                         whose extra output column is the bias gradient.  Layer 1's dX is the fp32 gradient handed to the
                         embedding backward; its dW may run on the side stream beside it."""

    @staticmethod
    def forward(ctx, x, ps, side_dw, *params):
        n_layers = (len(params) - 2) // 2
        ws, bs = params[0:2 * n_layers:2], params[1:2 * n_layers:2]
        w_out, b_out = params[-2], params[-1]
        planes = [P.split(x, ones_col=True)]
        masks = []
        for i in range(n_layers - 1):
            seed, off = _dropout_stream(x.shape[0] * ws[i].shape[0])
            if FUSE_FORWARD_EPILOGUE:
                yp, mask = P.linear_relu_dropout(planes[-1], weight_planes(ws[i]), bs[i], ps[i], seed, off,
                                                 _DROPOUT_DEV_COUNTER, ones_col=True)
            else:
                z = _fwd_gemm(planes[-1], ws[i], bs[i])
                yp, mask = P.relu_dropout_planes(z, ps[i], seed, off, _DROPOUT_DEV_COUNTER, ones_col=True)
            planes.append(yp)
            masks.append(mask)
        z = _fwd_gemm(planes[-1], ws[-1], bs[-1])
        y, mask, out = _relu_dropout_dot_fwd(z, ps[-1], w_out.reshape(-1), b_out)
        masks.append(mask)
        ctx.planes, ctx.masks, ctx.ps, ctx.side_dw, ctx.n_layers = planes, masks, ps, side_dw, n_layers
        ctx.save_for_backward(y, *params)
        return out.unsqueeze(1)

    @staticmethod
    def backward(ctx, g_out):
        y, *params = ctx.saved_tensors
        n_layers, ps, planes, masks = ctx.n_layers, ctx.ps, ctx.planes, ctx.masks
        ctx.planes = ctx.masks = None
        ws, bs = params[0:2 * n_layers:2], params[1:2 * n_layers:2]
        w_out, b_out = params[-2], params[-1]
        need = ctx.needs_input_grad                      # (x, ps, side_dw, W1, b1, ..., w_out, b_out)
        g = g_out.reshape(-1).contiguous()
        grads = [None] * len(params)
        if need[3 + 2 * n_layers]:
            grads[-2] = _colsum_weighted(y, g).reshape(w_out.shape)
        if b_out is not None and need[4 + 2 * n_layers]:
            grads[-1] = g.sum().reshape(1)
        gp = P.rank1_mask_planes(g, w_out.reshape(-1), masks[-1], ps[-1])
        gx = None
        for i in reversed(range(n_layers)):
            gp_prev = None
            if i > 0:
                gp_prev = P.dx_masked(gp, weight_planes(ws[i], transpose=True), masks[i - 1], ps[i - 1])
            elif need[0]:
                gx = _dx_gemm(gp, ws[0])
            want_w, want_b = need[3 + 2 * i], bs[i] is not None and need[4 + 2 * i]
            if want_w or want_b:
                if i == 0 and ctx.side_dw and gx is not None and _side_dw_safe(ws[0]) and \
                        (bs[0] is None or _side_dw_safe(bs[0])):
                    dw, db = _on_side_stream(lambda gp=gp, xp=planes[0]: P.gemm_dw(gp, xp, want_b), (gp, planes[0]))
                else:
                    dw, db = P.gemm_dw(gp, planes[i], want_b)
                grads[2 * i] = dw if want_w else None
                grads[2 * i + 1] = db if want_b else None
            gp = gp_prev
        return (gx, None, None, *grads)


class _MlpBatchNorm(torch.autograd.Function):
    """The dense tail  [Linear -> BatchNorm1d -> ReLU -> Dropout] x L  ->  Linear(hidden, 1)  in training mode
    (src/models/deepfm.py:55-66 with use_batchnorm, src/models/dcn.py:56-66) as ONE autograd node:

      fwd  layer i : tensor-core GEMM -> fp32 z; batch statistics (one pass over z, usually still in L2) + running
                     buffers; then ONE pass  z -> BatchNorm affine -> ReLU -> dropout -> planes of the next GEMM + mask.
                     Last layer: the same pass also forms the one-output Linear's dot product.
      bwd  layer i : the gradient w.r.t. the BatchNorm output arrives as fp32 (the next layer's dX GEMM applies the ReLU /
                     dropout mask in its epilogue; the head forms g[r] * w_out[c] * mask on the fly); one pass for
                     d gamma, d beta, one pass writes gz as planes; dW (+ bias gradient as its ones column) and dX GEMMs.
    torch's BatchNorm / threshold / dropout / sum kernels and every fp32 -> planes split of the layer-by-layer path are
    gone; what is saved for the backward is z (fp32), the planes of each layer's input, the masks and 4 N statistics."""

    @staticmethod
    def forward(ctx, x, ps, side_dw, bns, x_amax_slots, *params):
        n_layers = (len(params) - 2) // 4
        ws, bs = params[0:4 * n_layers:4], params[1:4 * n_layers:4]
        gammas, betas = params[2:4 * n_layers:4], params[3:4 * n_layers:4]
        w_out, b_out = params[-2], params[-1]
        fmt = MLP_PLANES_FORMAT
        # FP16X2: one zeroed device scalar per activation whose planes are written (the bounds their scales derive from)
        amaxs = torch.zeros(n_layers, 1, dtype=torch.float32, device=x.device) if fmt == P.FP16X2 else None
        x_amax = None
        if amaxs is not None:
            # max |x|: from the gather kernel's per-warp maxima when it supplied them, else one pass over x
            x_amax = P.absmax(x_amax_slots.view(1, -1) if x_amax_slots is not None else x, amaxs[0])
        planes = [P.split(x, ones_col=True, fmt=fmt, amax=x_amax)]
        masks, zs, stats = [], [], []
        y = out = None
        for i in range(n_layers):
            bn = bns[i]
            act_amax = amaxs[i + 1] if (amaxs is not None and i < n_layers - 1) else None
            z, st, affine = P.gemm_bn_stats(planes[-1], weight_planes(ws[i], fmt=planes[-1].fmt), bs[i], gammas[i],
                                            betas[i], bn.eps, bn.momentum,
                                            bn.running_mean if bn.track_running_stats else None,
                                            bn.running_var if bn.track_running_stats else None,
                                            act_amax=act_amax, bound_mul=1.0 / (1.0 - ps[i]), in_epilogue=BN_STATS_IN_GEMM)
            if bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            if i < n_layers - 1:
                seed, off = _dropout_stream(z.numel())
                yp, mask = P.bn_relu_dropout_planes(z, affine, ps[i], seed, off, _DROPOUT_DEV_COUNTER, ones_col=True,
                                                    fmt=fmt, amax=act_amax)
                planes.append(yp)
            else:
                y, mask, out = _relu_dropout_dot_fwd(z, ps[i], w_out.reshape(-1), b_out, affine)
            masks.append(mask)
            zs.append(z)
            stats.append(st)
        ctx.planes, ctx.masks, ctx.ps, ctx.side_dw, ctx.n_layers = planes, masks, ps, side_dw, n_layers
        ctx.n_z = len(zs)
        ctx.save_for_backward(y, *zs, *stats, *params)
        return out.unsqueeze(1)

    @staticmethod
    def backward(ctx, g_out):
        n_layers, ps, planes, masks = ctx.n_layers, ctx.ps, ctx.planes, ctx.masks
        ctx.planes = ctx.masks = None
        saved = ctx.saved_tensors
        y, zs, stats, params = saved[0], saved[1:1 + n_layers], saved[1 + n_layers:1 + 2 * n_layers], saved[1 + 2 * n_layers:]
        ws, bs = params[0:4 * n_layers:4], params[1:4 * n_layers:4]
        gammas, betas = params[2:4 * n_layers:4], params[3:4 * n_layers:4]
        w_out, b_out = params[-2], params[-1]
        need = ctx.needs_input_grad          # (x, ps, side_dw, bns, x_amax_slots, W1, b1, gamma1, beta1, ..., w_out, b_out)
        g = g_out.reshape(-1).contiguous()
        grads = [None] * len(params)
        if need[5 + 4 * n_layers]:
            grads[-2] = _colsum_weighted(y, g).reshape(w_out.shape)
        if b_out is not None and need[6 + 4 * n_layers]:
            grads[-1] = g.sum().reshape(1)
        g_r, _ = _relu_dropout_bwd_rank1(g, w_out.reshape(-1), masks[-1], ps[-1], False)   # fp32 [M, H]
        fmt = planes[0].fmt
        # FP16X2: max |g_r| per layer (zeroed device scalars): the head's from its two factors, the others from the
        # epilogue of the dX GEMM that produces them
        g_amaxs = torch.zeros(n_layers, 1, dtype=torch.float32, device=g.device) if fmt == P.FP16X2 else None
        if g_amaxs is not None:
            P.rank1_absmax(g, w_out.reshape(-1), ps[-1], g_amaxs[n_layers - 1])
        gx = None
        for i in reversed(range(n_layers)):
            gp, d_beta, d_gamma = P.bn_train_bwd_planes(g_r, zs[i], stats[i], gammas[i], fmt=fmt,
                                                        g_amax=g_amaxs[i] if g_amaxs is not None else None)
            if gammas[i] is not None and need[7 + 4 * i]:
                grads[4 * i + 2] = d_gamma
            if betas[i] is not None and need[8 + 4 * i]:
                grads[4 * i + 3] = d_beta
            g_prev = None
            if i > 0:
                g_prev = P.dx_masked(gp, weight_planes(ws[i], transpose=True, fmt=gp.fmt), masks[i - 1], ps[i - 1],
                                     to_planes=False, d_amax=g_amaxs[i - 1] if g_amaxs is not None else None)
            elif need[0]:
                gx = _dx_gemm(gp, ws[0])
            want_w, want_b = need[5 + 4 * i], bs[i] is not None and need[6 + 4 * i]
            if want_w or want_b:
                if i == 0 and ctx.side_dw and gx is not None and _side_dw_safe(ws[0]) and \
                        (bs[0] is None or _side_dw_safe(bs[0])):
                    dw, db = _on_side_stream(lambda gp=gp, xp=planes[0]: P.gemm_dw(gp, xp, want_b), (gp, planes[0]))
                else:
                    dw, db = P.gemm_dw(gp, planes[i], want_b)
                grads[4 * i] = dw if want_w else None
                grads[4 * i + 1] = db if want_b else None
            g_r = g_prev
        return (gx, None, None, None, None, *grads)


def _mlp_batchnorm_pattern(mods, x: torch.Tensor):
    """[Linear, BatchNorm1d, ReLU, Dropout] x L + Linear(h, 1), training mode, shapes the kernels take?"""
    if (len(mods) - 1) % 4 or len(mods) < 5 or not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        return None
    lins, bns, drops = [], [], []
    width = x.shape[1]
    for j in range(0, len(mods) - 1, 4):
        lin, bn, act, drop = mods[j], mods[j + 1], mods[j + 2], mods[j + 3]
        if not (isinstance(lin, torch.nn.Linear) and isinstance(bn, torch.nn.BatchNorm1d) and
                isinstance(act, torch.nn.ReLU) and isinstance(drop, torch.nn.Dropout)):
            return None
        h = lin.weight.shape[0]
        if lin.weight.shape[1] != width or h % 8 or width % 4 or not 0.0 <= drop.p < 1.0 or h > 2048 or \
                lin.weight.dtype != torch.float32 or not lin.weight.is_cuda or bn.num_features != h or \
                bn.momentum is None or not bn.training or (bn.weight is None) != (bn.bias is None) or \
                (bn.weight is not None and bn.weight.dtype != torch.float32):
            return None
        width = h
        lins.append(lin)
        bns.append(bn)
        drops.append(drop)
    head = mods[-1]
    if not _is_head(head, width) or x.shape[0] < 256:
        return None
    return lins, bns, drops, head


def _on_side_stream(fn, keep):
    """Run `fn` on the side stream (see _dw_on_side_stream); `keep` = main-pool tensors it reads."""
    dev = torch.cuda.current_device()
    main = torch.cuda.current_stream(dev)
    side = RF.side_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        out = fn()
    try:
        torch.autograd.Variable._execution_engine.queue_callback(lambda keep=keep: main.wait_stream(side))
    except Exception:  # noqa: BLE001 - not inside an autograd pass: join right away
        main.wait_stream(side)
    return out


def _mlp_relu_dropout_pattern(mods, x: torch.Tensor):
    """[Linear, ReLU, Dropout] x L + Linear(h, 1) with widths the plane kernels take (multiples of 8)?  Returns the
    (linears, dropouts, head) or None."""
    if (len(mods) - 1) % 3 or len(mods) < 4 or not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        return None
    lins, drops = [], []
    width = x.shape[1]
    for j in range(0, len(mods) - 1, 3):
        lin, act, drop = mods[j], mods[j + 1], mods[j + 2]
        if not (isinstance(lin, torch.nn.Linear) and isinstance(act, torch.nn.ReLU) and isinstance(drop, torch.nn.Dropout)):
            return None
        if lin.weight.shape[1] != width or lin.weight.shape[0] % 8 or width % 4 or not 0.0 < drop.p < 1.0 or \
                lin.weight.dtype != torch.float32 or not lin.weight.is_cuda or lin.weight.shape[0] > 2048:
            return None
        width = lin.weight.shape[0]
        lins.append(lin)
        drops.append(drop)
    head = mods[-1]
    if not _is_head(head, width) or x.shape[0] < 256:
        return None
    return lins, drops, head


def _is_head(mod, width: int) -> bool:
    """A Linear with ONE output on `width` inputs whose weight the glue kernels can read (fp32, 16-byte aligned)."""
    return (isinstance(mod, torch.nn.Linear) and mod.weight.shape[0] == 1 and mod.weight.shape[1] == width
            and width % 4 == 0 and mod.weight.is_cuda and mod.weight.dtype == torch.float32
            and mod.weight.data_ptr() % 16 == 0)


def run_sequential(seq: torch.nn.Sequential, x: torch.Tensor, overlap_first_dw: bool = False,
                   x_amax_slots: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Sequential forward of the dense tails with the same parameters / state dict, but
    Linear -> tensor-core GEMM, and (Linear ->) ReLU -> Dropout fused into single passes.
    BatchNorm1d (and anything else) runs unchanged.  overlap_first_dw: the input comes straight from the
    embedding gather, so the first layer's weight-gradient GEMM may run beside the embedding backward.
    x_amax_slots: device floats whose maximum is max |x| (the gather kernels fill them, functional.EMIT_AMAX)."""
    mods = list(seq)
    i = 0
    training = seq.training
    if training:
        pat = _mlp_batchnorm_pattern(mods, x)
        if pat is not None:
            lins, bns, drops, head = pat
            params = []
            for lin, bn in zip(lins, bns):
                params += [lin.weight, lin.bias, bn.weight, bn.bias]
            side = overlap_first_dw and not _has_hooks(lins[0].weight) and \
                (lins[0].bias is None or not _has_hooks(lins[0].bias))
            return _MlpBatchNorm.apply(x.contiguous(), tuple(float(d.p) for d in drops), side, tuple(bns), x_amax_slots,
                                       *params, head.weight, head.bias)
        pat = _mlp_relu_dropout_pattern(mods, x)
        if pat is not None:
            lins, drops, head = pat
            params = []
            for lin in lins:
                params += [lin.weight, lin.bias]
            side = overlap_first_dw and not _has_hooks(lins[0].weight) and \
                (lins[0].bias is None or not _has_hooks(lins[0].bias))
            return _MlpReluDropout.apply(x.contiguous(), tuple(float(d.p) for d in drops), side, *params, head.weight,
                                         head.bias)
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        nx2 = mods[i + 2] if i + 2 < len(mods) else None
        nx3 = mods[i + 3] if i + 3 < len(mods) else None
        if isinstance(m, torch.nn.Linear):
            ok = x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and \
                _use_kernel(x.shape[0], m.weight.shape[0], m.weight.shape[1], x, m.weight, m.bias)
            if ok and isinstance(nxt, torch.nn.ReLU) and isinstance(nx2, torch.nn.Dropout) and training and \
                    0.0 < nx2.p < 1.0 and m.weight.shape[0] <= 2048 and _is_head(nx3, m.weight.shape[0]) and \
                    i + 4 == len(mods):
                return _HeadBlock.apply(x, m.weight, m.bias, float(nx2.p), nx3.weight, nx3.bias)
            if ok and isinstance(nxt, torch.nn.ReLU) and isinstance(nx2, torch.nn.Dropout) and training and \
                    0.0 < nx2.p < 1.0 and m.weight.shape[0] <= 2048:
                x = _LinearReluDropout.apply(x, m.weight, m.bias, float(nx2.p),
                                             overlap_first_dw and i == 0 and not _has_hooks(m.weight))
                i += 3
                continue
            x = linear(x, m.weight, m.bias, side_dw=overlap_first_dw and i == 0 and not _has_hooks(m.weight))
            i += 1
            continue
        if isinstance(m, torch.nn.ReLU) and isinstance(nxt, torch.nn.Dropout) and training and 0.0 < nxt.p < 1.0 \
                and i + 3 == len(mods) and _fusable(x) and _is_head(nx2, x.shape[1]) and x.shape[0] >= 256:
            return _HeadBlock.apply(x, None, None, float(nxt.p), nx2.weight, nx2.bias)
        if isinstance(m, torch.nn.ReLU) and isinstance(nxt, torch.nn.Dropout):
            x = relu_dropout(x, float(nxt.p), training)
            i += 2
            continue
        x = m(x)
        i += 1
    return x
