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

from . import linalg as LA


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
            x_l = forward_mixture_layer(x_0, x_l, self.V[layer], self.C[layer], self.U[layer], self.biases[layer],
                                        self.gates, self._gate_softmax)
        return x_l
