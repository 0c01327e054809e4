"""DCN-Mix cross head behind the reference's module API (src/models/layer_dcn.py:27-115).

Same parameters (U/C/V ParameterLists, biases, gates), same math:
    H1 = tanh(x_l V_e); H2 = tanh(H1 C_e); Eo = H2 U_e
    x_{l+1} = sum_e (x_l . gates_e) * (Eo_e + b_l) * x_0 + x_l
restructured so the [B,E,Dm] expert outputs are never materialised: with
G2[b, e*r+k] = gate[b,e] * H2[b,e,k] the mixture is ONE GEMM against U viewed as
[E*r, Dm], and the (+b)*x_0 + x_l epilogue is folded around it.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import functional as RF
from . import linalg as LA
from . import planes as P


def forward_mixture_layer(x0, xl, V, C, U, bias, gates, gate_softmax: bool):
    """One cross layer. x0, xl [B,Dm]; V [E,Dm,r]; C [E,r,r]; U [E,r,Dm]; bias [1,Dm]; gates [E,Dm,1].

    The three genuine GEMMs (x_l V, H1 C block-diagonal, G2 U) run on the fp32-accurate
    tensor-core kernel (linalg.py); the [B,E] gate product (N = E = 4) stays a library call."""
    e, dm, r = V.shape
    bsz = xl.shape[0]
    vcat = V.permute(1, 0, 2).reshape(dm, e * r)           # [Dm, E*r]: column e*r+k = V[e,:,k]
    h1 = torch.tanh(LA.matmul(xl, vcat))                   # [B, E*r]   (layer_dcn.py:20-21)
    h2 = torch.tanh(LA.expert_matmul(h1, C))               # [B, E*r]   (:22)
    g = torch.matmul(xl, gates.squeeze(2).t())             # [B, E]     (:107-109)
    if gate_softmax:
        g = torch.softmax(g, dim=1)
    g2 = (h2.view(bsz, e, r) * g.unsqueeze(2)).reshape(bsz, e * r)
    t = LA.matmul(g2, U.reshape(e * r, dm)) + g.sum(1, keepdim=True) * bias   # sum_e g_e (Eo_e + b)  (:23,102)
    return x0 * t + xl                                     # (:103,113)


class _CrossLayer(torch.autograd.Function):
    """One DCN-Mix cross layer with every element-wise step fused into four custom passes around the
    three tensor-core GEMMs (csrc/dcn.cu); identity gate.  Every GEMM operand is split into bf16 planes once and
    the planes are shared by all GEMMs that read it (x_l: forward, d_gates, dV; H1: expert GEMM, dC; G2: mixture
    GEMM, dU; ...); the experts are a batch addressed as column blocks inside the stored matrices."""

    @staticmethod
    def forward(ctx, x0, xl, V, C, U, bias, gates):
        lib = L.load()
        e, dm, r = V.shape
        bsz = xl.shape[0]
        er = e * r
        dev = xl.device
        x0 = x0.contiguous()
        xl = xl.contiguous()
        st = L.stream_ptr(dev)
        vcat_p = P.split(V.detach().permute(1, 0, 2).reshape(dm, er).contiguous())   # stored [Dm, E*r]
        cc_p = P.split(C.detach().reshape(er, r))                                        # stored [E*r, r]
        ucat_p = P.split(U.detach().reshape(er, dm))                                     # stored [E*r, Dm]
        g2d = gates.detach().reshape(e, dm)
        b1d = bias.detach().reshape(dm)
        xl_p = P.split(xl)
        h1 = P.gemm(xl_p, vcat_p, bsz, er, dm, b_mn_major=True, split_k=1).tanh_()     # x_l V_cat (layer_dcn.py:20-21)
        h1_p = P.split(h1)
        p2 = torch.empty(bsz, er, dtype=torch.float32, device=dev)
        # P2[:, l, :] = H1[:, l, :] @ C[l]  (:22): A column step r, B = C[l] stored [K = r, N = r] at row l*r
        P.gemm(h1_p, cc_p, bsz, r, r, b_mn_major=True, out=p2, batch=e, a_steps=(0, r), b_steps=(r, 0),
               d_batch_stride=r, split_k=1)
        g = torch.empty(bsz, e, dtype=torch.float32, device=dev)
        g2 = torch.empty(bsz, er, dtype=torch.float32, device=dev)
        RF._call("dcn_gate_mix_fwd", lib.rsb_dcn_gate_mix_fwd, L.ptr(p2), L.ptr(xl), L.ptr(g2d), bsz, dm, e, r,
                 L.ptr(p2), L.ptr(g), L.ptr(g2), st, nbytes=bsz * (dm + 3 * er) * 4)     # H2 overwrites P2
        h2 = p2
        g2_p = P.split(g2)
        t0 = P.gemm(g2_p, ucat_p, bsz, dm, er, b_mn_major=True, split_k=1)             # G2 U_cat (:23)
        out = torch.empty(bsz, dm, dtype=torch.float32, device=dev)
        RF._call("dcn_cross_out_fwd", lib.rsb_dcn_cross_out_fwd, L.ptr(t0), L.ptr(x0), L.ptr(xl), L.ptr(b1d), L.ptr(g),
                 bsz, dm, e, L.ptr(out), st, nbytes=bsz * dm * 16)
        ctx.save_for_backward(x0, xl, b1d, g2d, h1, h2, g, t0)
        ctx.planes = (xl_p, h1_p, g2_p, vcat_p, cc_p, ucat_p)
        ctx.dims = (e, dm, r)
        return out

    @staticmethod
    def backward(ctx, g_next):
        lib = L.load()
        x0, xl, b1d, g2d, h1, h2, g, t0 = ctx.saved_tensors
        xl_p, h1_p, g2_p, vcat_p, cc_p, ucat_p = ctx.planes
        ctx.planes = None
        e, dm, r = ctx.dims
        er = e * r
        bsz = xl.shape[0]
        dev = xl.device
        st = L.stream_ptr(dev)
        g_next = g_next.contiguous()
        gT = torch.empty(bsz, dm, dtype=torch.float32, device=dev)
        gx0 = torch.empty(bsz, dm, dtype=torch.float32, device=dev)
        dsg = torch.empty(bsz, dtype=torch.float32, device=dev)
        RF._call("dcn_cross_out_bwd", lib.rsb_dcn_cross_out_bwd, L.ptr(g_next), L.ptr(x0), L.ptr(t0), L.ptr(b1d),
                 L.ptr(g), bsz, dm, e, L.ptr(gT), L.ptr(gx0), L.ptr(dsg), st, nbytes=bsz * dm * 20)
        d_bias = torch.mv(gT.t(), g.sum(1)).reshape(1, dm)           # sum_b gT[b,:] * sg[b]
        gT_p = P.split(gT)
        g_g2 = P.gemm(gT_p, ucat_p, bsz, er, dm, split_k=1)          # gT @ U_cat^T: U_cat stored [N = E*r, K = Dm]
        d_u = P.gemm(g2_p, gT_p, er, dm, bsz, a_mn_major=True, b_mn_major=True, split_k=0).reshape(e, r, dm)
        dg = torch.empty(bsz, e, dtype=torch.float32, device=dev)
        g_xl = torch.empty(bsz, dm, dtype=torch.float32, device=dev)
        RF._call("dcn_gate_mix_bwd", lib.rsb_dcn_gate_mix_bwd, L.ptr(g_g2), L.ptr(h2), L.ptr(g), L.ptr(dsg), L.ptr(g2d),
                 L.ptr(g_next), bsz, dm, e, r, L.ptr(g_g2), L.ptr(dg), L.ptr(g_xl), st,
                 nbytes=bsz * (3 * er + 2 * dm) * 4)                 # gP2 overwrites gG2
        g_p2 = g_g2
        # [E,B] x [B,Dm] (M = E = 4 rows of one tile): split over the batch
        d_gates = P.gemm(P.split(dg), xl_p, e, dm, bsz, a_mn_major=True, b_mn_major=True, split_k=0).reshape(e, dm, 1)
        # block-diagonal expert GEMM backward
        g_p2_p = P.split(g_p2)
        g_h1 = torch.empty(bsz, er, dtype=torch.float32, device=dev)
        # gH1[:, l, :] = gP2[:, l, :] @ C[l]^T : C[l] read as stored [N = r, K = r] at row l*r
        P.gemm(g_p2_p, cc_p, bsz, r, r, out=g_h1, batch=e, a_steps=(0, r), b_steps=(r, 0), d_batch_stride=r, split_k=1)
        # dC[l] = H1[:, l, :]^T @ gP2[:, l, :] : batched split-K, both operands MN-major with column step r
        d_c = torch.empty(e, r, r, dtype=torch.float32, device=dev)
        P.gemm(h1_p, g_p2_p, r, r, bsz, a_mn_major=True, b_mn_major=True, out=d_c.view(er, r), batch=e,
               a_steps=(0, r), b_steps=(0, r), d_batch_stride=r * r, split_k=0)
        g_p1 = torch.ops.aten.tanh_backward(g_h1, h1)
        g_p1_p = P.split(g_p1)
        # g_xl += gP1 @ V_cat^T : V_cat stored [N = Dm, K = E*r]
        P.gemm(g_p1_p, vcat_p, bsz, dm, er, beta=1.0, c=g_xl, out=g_xl, split_k=1)
        d_vcat = P.gemm(xl_p, g_p1_p, dm, er, bsz, a_mn_major=True, b_mn_major=True, split_k=0)
        d_v = d_vcat.reshape(dm, e, r).permute(1, 0, 2).contiguous()
        return gx0, g_xl, d_v, d_c, d_u, d_bias, d_gates


def _fused_layer_ok(x0, xl, V, gate_softmax: bool) -> bool:
    e, dm, r = V.shape
    bsz = xl.shape[0]
    return (xl.is_cuda and not gate_softmax and dm % 4 == 0 and r % 32 == 0 and e <= 8 and bsz >= 512
            and xl.dtype == torch.float32)


class DCN_MixHead(nn.Module):
    def __init__(self, num_experts: int, num_layers: int, rank: int, hidden_size: int,
                 activation: Optional[str] = None, gate_act: str = "identity"):
        super().__init__()
        self.num_experts = num_experts
        self.num_layers = num_layers
        self.rank = rank
        assert gate_act in ["softmax", "identity"]
        self.U = nn.ParameterList([self._init_parameters((num_experts, rank, hidden_size)) for _ in range(num_layers)])
        self.C = nn.ParameterList([self._init_parameters((num_experts, rank, rank)) for _ in range(num_layers)])
        self.V = nn.ParameterList([self._init_parameters((num_experts, hidden_size, rank)) for _ in range(num_layers)])
        self.biases = nn.ParameterList([self._init_parameters((1, hidden_size), "zeros") for _ in range(num_layers)])
        self.gates = self._init_parameters((num_experts, hidden_size, 1))
        self.gate_act = nn.Softmax(dim=1) if gate_act == "softmax" else nn.Identity()
        self._gate_softmax = gate_act == "softmax"
        self.act_name = "tanh"  # the reference ignores `activation` too (layer_dcn.py:78-79)
        self.act = nn.Tanh()

    def _init_parameters(self, shape, dist="he") -> nn.Parameter:
        if dist == "zeros":
            return nn.Parameter(torch.zeros(*shape))
        tensor = torch.empty(*shape)
        nn.init.kaiming_normal_(tensor)
        return nn.Parameter(tensor)

    def forward(self, x_0):
        x_l = x_0
        for layer in range(self.num_layers):
            if _fused_layer_ok(x_0, x_l, self.V[layer], self._gate_softmax):
                x_l = _CrossLayer.apply(x_0, x_l, self.V[layer], self.C[layer], self.U[layer], self.biases[layer],
                                        self.gates)
            else:
                x_l = forward_mixture_layer(x_0, x_l, self.V[layer], self.C[layer], self.U[layer],
                                            self.biases[layer], self.gates, self._gate_softmax)
        return x_l
