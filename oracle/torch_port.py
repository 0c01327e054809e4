"""CPU baseline port of the reference's CTR train step (TEST / BENCH INFRASTRUCTURE ONLY).

The reference's CPU path for this benchmark IS PyTorch's CPU kernels driven from Python
(`F.embedding`, `nn.EmbeddingBag`, elementwise FM ops, `torch.optim.Adam/SparseAdam`;
src/models/deepfm.py:79-105,155-219, src/models/embeddings/qr_embedding.py:95-109,
src/models/layer_dcn.py:8-115, src/trainer/deepfm.py:44-62).  /root/reference does not
exist on the GPU box, so this file restates that path functionally (plain tensors in a
dict, the same torch CPU operators in the same order, autograd for the backward) so that
`bench.py` can time "the reference's CPU PyTorch path" next to the GPU number.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` (cpu_baseline / --impl reference)
may import it.  It is pinned against the golden vectors in tests/test_oracle_golden.py
(test_torch_port_*).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def make_deepfm_params(field_dims: List[int], d: int, hidden: List[int], emb_cfg: Dict, use_bn: bool,
                       seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init parameters with the reference's shapes and init distributions."""
    g = torch.Generator().manual_seed(seed)
    n = sum(field_dims)
    p: Dict[str, torch.Tensor] = {}
    name = emb_cfg.get("name", "vanilla")
    if name == "vanilla":
        a = math.sqrt(6.0 / (n + d))
        p["embedding._emb_module.weight"] = torch.empty(n, d).uniform_(-a, a, generator=g)
    elif name == "qr":
        div = emb_cfg.get("divider") or int(math.sqrt(n))
        alpha = math.sqrt(1 / n)
        p["embedding.emb1.weight"] = torch.empty(div, d).uniform_(alpha, 1, generator=g)
        p["embedding.emb2.weight"] = torch.empty((n - 1) // div + 1, d).uniform_(alpha, 1, generator=g)
    elif name == "pep":
        a = math.sqrt(6.0 / (n + d))
        p["embedding.emb.weight"] = torch.empty(n, d).uniform_(-a, a, generator=g)
        shape = {"global": (1,), "dimension": (d,), "feature": (n, 1), "feature_dim": (n, d)}[
            emb_cfg.get("threshold_type", "feature_dim")]
        p["embedding.s"] = torch.full(shape, float(emb_cfg.get("init_threshold", -150)))
    else:
        raise NotImplementedError(name)
    if emb_cfg.get("_no_first_order"):
        for v in p.values():
            v.requires_grad_(True)
        return p
    p["fc.weight"] = torch.randn(n, 1, generator=g)
    p["_bias"] = torch.zeros(1)
    inp = d * len(field_dims)
    li = 0
    for h in hidden:
        k = 1 / math.sqrt(inp)
        p[f"_deep_branch.{li}.weight"] = torch.empty(h, inp).uniform_(-k, k, generator=g)
        p[f"_deep_branch.{li}.bias"] = torch.empty(h).uniform_(-k, k, generator=g)
        li += 1
        if use_bn:
            p[f"_deep_branch.{li}.weight"] = torch.ones(h)
            p[f"_deep_branch.{li}.bias"] = torch.zeros(h)
            li += 1
        li += 2  # ReLU, Dropout
        inp = h
    k = 1 / math.sqrt(inp)
    p[f"_deep_branch.{li}.weight"] = torch.empty(1, inp).uniform_(-k, k, generator=g)
    p[f"_deep_branch.{li}.bias"] = torch.empty(1).uniform_(-k, k, generator=g)
    for v in p.values():
        v.requires_grad_(True)
    return p


def embedding_forward(p: Dict[str, torch.Tensor], rows: torch.Tensor, emb_cfg: Dict) -> torch.Tensor:
    name = emb_cfg.get("name", "vanilla")
    if name == "vanilla":
        return F.embedding(rows, p["embedding._emb_module.weight"], sparse=bool(emb_cfg.get("sparse", False)))
    if name == "pep":       # src/models/embeddings/pep_embedding.py:82-92
        v = p["embedding.emb.weight"]
        return F.embedding(rows, torch.sign(v) * torch.relu(torch.abs(v) - torch.sigmoid(p["embedding.s"])))
    div = p["embedding.emb1.weight"].shape[0]
    e1 = F.embedding(rows % div, p["embedding.emb1.weight"])
    e2 = F.embedding(rows // div, p["embedding.emb2.weight"])
    op = emb_cfg.get("operation", "mult")
    if op == "mult":
        return e1 * e2
    if op == "add":
        return e1 + e2
    return torch.cat([e1, e2], dim=1)


def mlp_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str, p_dropout: float, training: bool,
                bn_state: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in p if k.startswith(prefix)})
    lin = [i for i in idx if p[f"{prefix}{i}.weight"].dim() == 2]
    for j, i in enumerate(lin):
        x = F.linear(x, p[f"{prefix}{i}.weight"], p[f"{prefix}{i}.bias"])
        if j == len(lin) - 1:
            break
        if f"{prefix}{i + 1}.weight" in p and p[f"{prefix}{i + 1}.weight"].dim() == 1:
            rm = rv = None
            if bn_state is not None:
                rm, rv = bn_state[f"{prefix}{i + 1}.running_mean"], bn_state[f"{prefix}{i + 1}.running_var"]
            x = F.batch_norm(x, rm, rv, p[f"{prefix}{i + 1}.weight"], p[f"{prefix}{i + 1}.bias"],
                             training=training or rm is None)
        x = F.relu(x)
        x = F.dropout(x, p_dropout, training)
    return x


def deepfm_forward(p, x, offsets, emb_cfg, p_dropout=0.0, training=True, bn_state=None):
    """src/models/deepfm.py:79-105."""
    rows = x + offsets
    emb = embedding_forward(p, rows, emb_cfg)
    square_of_sum = emb.sum(dim=1).pow(2)
    sum_of_square = emb.pow(2).sum(dim=1)
    first = F.embedding_bag(rows, p["fc.weight"], mode="sum") + p["_bias"]
    y_fm = first + 0.5 * (square_of_sum - sum_of_square).sum(1, keepdim=True)
    b = emb.shape[0]
    deep = mlp_forward(p, emb.reshape(b, -1), "_deep_branch.", p_dropout, training, bn_state)
    return (y_fm + deep).squeeze(-1)


def make_dcn_params(field_dims: List[int], d: int, hidden: List[int], emb_cfg: Dict, num_layers: int = 3,
                    num_experts: int = 4, rank: int = 64, seed: int = 0) -> Dict[str, torch.Tensor]:
    """DCN_Mix parameters with the reference's shapes / init (src/models/dcn.py:11-74, layer_dcn.py:46-88)."""
    cfg = dict(emb_cfg)
    cfg["_no_first_order"] = True
    p = make_deepfm_params(field_dims, d, [], cfg, False, seed)
    g = torch.Generator().manual_seed(seed + 1)
    dm = d * len(field_dims)

    def he(*shape):
        fan_in = shape[1] * (shape[2] if len(shape) > 2 else 1)          # torch's fan_in for a 3-D tensor
        return torch.randn(*shape, generator=g) * math.sqrt(2.0 / fan_in)

    for l in range(num_layers):
        p[f"cross_head.U.{l}"] = he(num_experts, rank, dm)
        p[f"cross_head.C.{l}"] = he(num_experts, rank, rank)
        p[f"cross_head.V.{l}"] = he(num_experts, dm, rank)
        p[f"cross_head.biases.{l}"] = torch.zeros(1, dm)
    p["cross_head.gates"] = he(num_experts, dm, 1)
    inp, li = dm, 0
    for h in hidden:
        k = 1 / math.sqrt(inp)
        p[f"_dnn.{li}.weight"] = torch.empty(h, inp).uniform_(-k, k, generator=g)
        p[f"_dnn.{li}.bias"] = torch.empty(h).uniform_(-k, k, generator=g)
        p[f"_dnn.{li + 1}.weight"] = torch.ones(h)
        p[f"_dnn.{li + 1}.bias"] = torch.zeros(h)
        li += 4
        inp = h
    k = 1 / math.sqrt(inp)
    p[f"_dnn.{li}.weight"] = torch.empty(1, inp).uniform_(-k, k, generator=g)
    p[f"_dnn.{li}.bias"] = torch.empty(1).uniform_(-k, k, generator=g)
    for v in p.values():
        v.requires_grad_(True)
    return p


def dcn_mix_forward(p, x, offsets, emb_cfg, p_dropout=0.0, training=True, bn_state=None):
    """src/models/dcn.py:76-96 + layer_dcn.py:8-24,90-115, the reference's operators in its order
    (`x @ V`, permute, the two einsums as batched matmuls, tanh always, identity gate)."""
    rows = x + offsets
    emb = embedding_forward(p, rows, emb_cfg)
    x0 = emb.reshape(emb.shape[0], -1)
    xl = x0
    x0u = x0.unsqueeze(1)
    layers = sorted({int(k.split(".")[2]) for k in p if k.startswith("cross_head.U.")})
    for l in layers:
        C, U, V, b = (p[f"cross_head.{n}.{l}"] for n in ("C", "U", "V", "biases"))
        e = torch.tanh(xl @ V).permute(1, 0, 2)                          # [B, E, r]
        e = torch.tanh(torch.einsum("ber,ers->bes", e, C))
        e = torch.einsum("ber,erd->bed", e, U)
        e = x0u * (e + b)
        gates = (xl @ p["cross_head.gates"]).squeeze(2).permute(1, 0)     # [B, E]
        xl = torch.einsum("be,bed->bd", gates, e) + xl
    return mlp_forward(p, xl, "_dnn.", p_dropout, training, bn_state).squeeze(-1)


def make_optimizers(p: Dict[str, torch.Tensor], cfg: Dict):
    """src/models/deepfm.py:155-219."""
    sparse = cfg.get("sparse", False)
    if sparse:
        emb = [v for k, v in p.items() if k.startswith("embedding.")]
        rest = [v for k, v in p.items() if not k.startswith("embedding.")]
        return [torch.optim.SparseAdam(emb, lr=cfg.get("learning_rate_emb", cfg["learning_rate"])),
                torch.optim.Adam(rest, lr=cfg["learning_rate"], weight_decay=cfg["weight_decay"])]
    return [torch.optim.Adam(list(p.values()), lr=cfg["learning_rate"], weight_decay=cfg["weight_decay"])]


def train_step(p, opts, x, y, offsets, emb_cfg, p_dropout, sync: bool = True, forward=None):
    """One iteration of src/trainer/deepfm.py:44-62 (sync=False skips the loss.item() read-back)."""
    logits = (forward or deepfm_forward)(p, x, offsets, emb_cfg, p_dropout, training=True)
    loss = F.binary_cross_entropy_with_logits(logits, y.float())
    for o in opts:
        o.zero_grad()
    loss.backward()
    for o in opts:
        o.step()
    return loss.item() if sync else loss
