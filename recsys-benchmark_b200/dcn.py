"""DCN-Mix behind the reference's model API (src/models/dcn.py:11-129)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union, cast

import torch
from torch import nn

from .embeddings import IEmbedding
from .layer_dcn import DCN_MixHead
from .linalg import run_sequential


class DCN_Mix(nn.Module):
    embedding: IEmbedding

    def __init__(self, field_dims: List[int], num_factor: int, hidden_sizes: List[int], num_layers: int = 3,
                 num_experts: int = 4, rank: int = 64, activation: Optional[str] = None,
                 embedding_config: Optional[Dict] = None, p_dropout=0.5, empty_embedding=False):
        super().__init__()
        from . import get_embedding

        if not embedding_config:
            embedding_config = {"name": "vanilla"}
        if not empty_embedding:
            self.embedding = get_embedding(embedding_config, field_dims, num_factor, mode=None, field_name="dcn")
        inp_size = num_factor * len(field_dims)
        self.cross_head = DCN_MixHead(num_experts, num_layers, rank, inp_size, activation)
        layers: List[nn.Module] = []
        for size in hidden_sizes:
            layers.append(nn.Linear(inp_size, size))
            layers.append(nn.BatchNorm1d(size))
            layers.append(nn.ReLU())
            layers.append(nn.Dropout(p_dropout))
            inp_size = size
        layers.append(nn.Linear(inp_size, 1))
        self._dnn = nn.Sequential(*layers)
        dims = torch.cat([torch.tensor([0], dtype=torch.long), torch.tensor(field_dims)])
        self.register_buffer("offsets", torch.cumsum(dims[:-1], 0).unsqueeze(0))

    def forward(self, x):
        """x: [B, F] per-field ids without offsets -> logits [B] (src/models/dcn.py:76-96)."""
        emb, _ = self.embedding.lookup(x, self.offsets)   # offsets add fused into the gather
        cross = self.cross_head(emb.reshape(emb.shape[0], emb.shape[1] * emb.shape[2]))
        return run_sequential(self._dnn, cross).squeeze(-1)

    @classmethod
    def load(cls, checkpoint: Union[str, Dict[str, Any]], strict=True, *, empty_embedding=False):
        if isinstance(checkpoint, str):
            checkpoint = torch.load(checkpoint, map_location="cpu")
        checkpoint = cast(Dict[str, Any], checkpoint)
        model_config = dict(checkpoint["model_config"])
        model_config.pop("compile_model", None)  # custom kernels are opaque to Dynamo: never compiled
        model = cls(checkpoint["field_dims"], **model_config, empty_embedding=empty_embedding)
        state = {k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k: v
                 for k, v in checkpoint["state_dict"].items()}
        model.load_state_dict(state, strict=strict)
        return model
