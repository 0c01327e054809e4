"""DeepFM behind the reference's model API (src/models/deepfm.py:11-219).

Same constructor, parameter / buffer names and shapes (state dicts are interchangeable
with the reference's), same `get_optimizers` / `save_model_checkpoint` helpers.  The
forward differs in HOW, not what: offsets add + embedding gather (+ variant) + first-order
EmbeddingBag + FM second order are ONE fused launch (`embedding.lookup`), and the
backward is the sorted segmented scatter-add.  The MLP stays in cuBLAS (off the hot path).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Union, cast

import torch
from torch import nn

from .embeddings import IEmbedding
from .linalg import run_sequential


class DeepFM(nn.Module):
    embedding: IEmbedding

    def __init__(self, field_dims: List[int], num_factor: int, hidden_sizes: List[int], p_dropout: float = 0.1,
                 use_batchnorm=False, embedding_config: Optional[Dict] = None, empty_embedding=False):
        super().__init__()
        from . import get_embedding

        if not embedding_config:
            embedding_config = {"name": "vanilla"}
        num_inputs = sum(field_dims)
        if not empty_embedding:
            self.embedding = get_embedding(embedding_config, field_dims, num_factor, mode=None, field_name="deepfm")
        self.fc = nn.EmbeddingBag(num_inputs, 1, mode="sum")
        self.linear_layer = nn.Linear(1, 1)  # present (unused) in the reference too: keeps state dicts equal
        self._bias = nn.Parameter(torch.zeros(1))

        deep_branch_inp = num_factor * len(field_dims)
        layers: List[nn.Module] = []
        for size in hidden_sizes:
            layers.append(nn.Linear(deep_branch_inp, size))
            if use_batchnorm:
                layers.append(nn.BatchNorm1d(size))
            layers.append(nn.ReLU())
            layers.append(nn.Dropout(p_dropout))
            deep_branch_inp = size
        layers.append(nn.Linear(deep_branch_inp, 1))
        self._deep_branch = nn.Sequential(*layers)

        dims = torch.cat([torch.tensor([0], dtype=torch.long), torch.tensor(field_dims)])
        self.register_buffer("offsets", torch.cumsum(dims[:-1], 0).unsqueeze(0))

    def forward(self, x):
        """x: [B, F] per-field ids (int32 or int64, WITHOUT offsets) -> logits [B]."""
        emb, y_fm = self.embedding.lookup(x, self.offsets, self.fc.weight, self._bias)
        b = emb.shape[0]
        scores = y_fm.unsqueeze(1) + run_sequential(self._deep_branch, emb.reshape(b, emb.shape[1] * emb.shape[2]),
                                                       overlap_first_dw=True,
                                                       x_amax_slots=getattr(emb, "_rsb_amax_slots", None))
        return scores.squeeze(-1)

    def get_ranks(self, x) -> torch.Tensor:
        return torch.argsort(self(x), descending=True)

    @classmethod
    def load(cls, checkpoint: Union[str, Dict[str, Any]], strict=True, *, empty_embedding=False) -> "DeepFM":
        if isinstance(checkpoint, str):
            checkpoint = torch.load(checkpoint, map_location="cpu")
        checkpoint = cast(Dict[str, Any], checkpoint)
        model = cls(checkpoint["field_dims"], **checkpoint["model_config"], empty_embedding=empty_embedding)
        model.load_state_dict(checkpoint["state_dict"], strict=strict)
        return model


def save_model_checkpoint(model: DeepFM, checkpoint_dir: str, name: str = "target"):
    """{checkpoint_dir}/deepfm/{name}.pth (src/models/deepfm.py:137-152)."""
    field_dir = os.path.join(checkpoint_dir, "deepfm")
    os.makedirs(field_dir, exist_ok=True)
    torch.save(model.embedding.state_dict(), os.path.join(field_dir, f"{name}.pth"))


def get_optimizers(model: nn.Module, config: Dict) -> List[torch.optim.Optimizer]:
    """Same grouping rules as src/models/deepfm.py:155-219.

    Extra opt-in key `fused_sparse: true` (with `sparse: true`): the embedding table is
    updated by the fused segmented-reduce + SparseAdam / SGD row kernel instead of
    torch.optim.SparseAdam / SGD on a COO gradient (same arithmetic, see optim.py)."""
    from .optim import FusedDenseAdam, FusedSparseAdam, FusedSparseSGD

    def dense_adam(params):
        # `fused_adam` (opt-in): "rsb" = the one-launch Adam of this library (optim.FusedDenseAdam), true = torch's
        # multi-tensor fused Adam; same arithmetic as the default torch.optim.Adam either way
        if config.get("fused_adam") == "rsb" and not config.get("capturable", False):
            return FusedDenseAdam(params, lr=config["learning_rate"], weight_decay=config["weight_decay"])
        return torch.optim.Adam(params, lr=config["learning_rate"], weight_decay=config["weight_decay"],
                                fused=bool(config.get("fused_adam", False)),
                                capturable=bool(config.get("capturable", False)))

    sparse: bool = config.get("sparse", False)
    optimizer_name: str = config.get("optimizer", "adam")
    lr_emb = config.get("learning_rate_emb", config["learning_rate"])
    fused = bool(config.get("fused_sparse", False))

    decay_param = []
    no_decay_param = []
    if sparse:
        decay_param = [p for name, p in model.named_parameters() if "embedding." not in name]
        no_decay_param = list(model.embedding.parameters())

    if sparse and optimizer_name == "adam":
        emb_opt = (FusedSparseAdam(model.embedding, lr=lr_emb) if fused
                   else torch.optim.SparseAdam(no_decay_param, lr=lr_emb))
        return [emb_opt, dense_adam(decay_param)]
    if optimizer_name == "adam":
        return [dense_adam(model.parameters())]
    if optimizer_name == "sgd":
        if not sparse:
            return [torch.optim.SGD(model.parameters(), lr=config["learning_rate"],
                                    weight_decay=config["weight_decay"])]
        if fused:
            return [FusedSparseSGD(model.embedding, lr=lr_emb),
                    torch.optim.SGD(decay_param, lr=config["learning_rate"], weight_decay=config["weight_decay"])]
        return [torch.optim.SGD(
            [dict(params=no_decay_param, weight_decay=0, lr=lr_emb),
             dict(params=decay_param, weight_decay=config["weight_decay"], lr=config["learning_rate"])],
            config["learning_rate"])]
    raise ValueError(f"{optimizer_name=} is not recognized")
