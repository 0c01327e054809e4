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
    else:
        raise NotImplementedError(name)
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


def make_optimizers(p: Dict[str, torch.Tensor], cfg: Dict):
    """src/models/deepfm.py:155-219."""
    sparse = cfg.get("sparse", False)
    if sparse:
        emb = [v for k, v in p.items() if k.startswith("embedding.")]
        rest = [v for k, v in p.items() if not k.startswith("embedding.")]
        return [torch.optim.SparseAdam(emb, lr=cfg.get("learning_rate_emb", cfg["learning_rate"])),
                torch.optim.Adam(rest, lr=cfg["learning_rate"], weight_decay=cfg["weight_decay"])]
    return [torch.optim.Adam(list(p.values()), lr=cfg["learning_rate"], weight_decay=cfg["weight_decay"])]


def train_step(p, opts, x, y, offsets, emb_cfg, p_dropout, sync: bool = True):
    """One iteration of src/trainer/deepfm.py:44-62 (sync=False skips the loss.item() read-back)."""
    logits = deepfm_forward(p, x, offsets, emb_cfg, p_dropout, training=True)
    loss = F.binary_cross_entropy_with_logits(logits, y.float())
    for o in opts:
        o.zero_grad()
    loss.backward()
    for o in opts:
        o.step()
    return loss.item() if sync else loss
