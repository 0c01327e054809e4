"""recsys-benchmark CTR hot path, B200-native (import as `recsys_benchmark_b200`).

Public surface mirrors the reference's plugin / model API for this path
(src/models/__init__.py:69-131, src/models/embeddings/__init__.py:18-73):
NAME_TO_CLS, get_embedding, get_ctr_model, load_ctr_model, save_ctr_checkpoint,
DeepFM, DCN_Mix, get_optimizers — plus `install_into_reference()` which rebinds the
reference's own registry to these classes so its trainers and scripts run unchanged.
"""
from __future__ import annotations

import copy
import os
from typing import Dict, List, Optional, Union

import torch

from . import _lib, embeddings
from .embeddings import (CerpEmbedding, IEmbedding, OptEmbed, RetrainCerpEmbedding, PepEmbeeding, QRHashingEmbedding, RetrainOptEmbed,
                         RetrainPepEmbedding, VanillaEmbedding)

__version__ = "0.1.0"

# Names are the reference's (src/models/embeddings/__init__.py:18-36); only the variants on
# the hot path are provided.
NAME_TO_CLS = {
    "vanilla": VanillaEmbedding,
    "qr": QRHashingEmbedding,
    "pep": PepEmbeeding,
    "pep_retrain": RetrainPepEmbedding,
    "deepfm_optembed": OptEmbed,
    "deepfm_optembed_d": OptEmbed,
    "deepfm_optembed_retrain": RetrainOptEmbed,
    "cerp": CerpEmbedding,
    "cerp_retrain": RetrainCerpEmbedding,
}


def get_embedding(embedding_config: Dict, field_dims: Union[int, List[int]], hidden_size: int,
                  mode: Optional[str] = None, field_name: str = "") -> IEmbedding:
    """Factory with the reference's semantics (src/models/embeddings/__init__.py:39-73)."""
    assert mode in [None, "sum", "mean", "max"], "Unsupported mode"
    name = embedding_config["name"]
    cfg = copy.deepcopy(embedding_config)
    cfg.pop("name")
    if name not in NAME_TO_CLS:
        raise NotImplementedError(f"{name} not found in mapping from name to class")
    if name.startswith("pep") or name.startswith("cerp"):
        cfg["field_name"] = field_name
    if name == "deepfm_optembed_d":
        cfg["t_init"] = None
    return NAME_TO_CLS[name](field_dims, hidden_size, mode=mode, **cfg)


from .dcn import DCN_Mix  # noqa: E402
from .deepfm import DeepFM, get_optimizers, save_model_checkpoint  # noqa: E402
from .layer_dcn import DCN_MixHead  # noqa: E402
from .optim import FusedSparseAdam, FusedSparseSGD  # noqa: E402
from .pruned import PrunedEmbedding  # noqa: E402
from .dhe import DHEmbedding  # noqa: E402

NAME_TO_CLS["dhe"] = DHEmbedding


def get_ctr_model(field_dims, model_config: dict):
    """src/models/__init__.py:69-89.  `compile_model` is accepted and ignored: the custom
    kernels are opaque to Dynamo, and there is nothing left for Inductor to fuse."""
    name = model_config.pop("name") if "name" in model_config else "deepfm"
    if name == "deepfm":
        return DeepFM(field_dims, **model_config)
    if name == "dcn_mix":
        compile_model = model_config.pop("compile_model", True)
        model = DCN_Mix(field_dims, **model_config)
        model_config["compile_model"] = compile_model
        return model
    raise NotImplementedError()


def load_ctr_model(model_config, checkpoint, strict=True, *, empty_embedding=False):
    name = model_config.pop("name") if "name" in model_config else "deepfm"
    if name == "deepfm":
        return DeepFM.load(checkpoint, strict, empty_embedding=empty_embedding)
    if name == "dcn_mix":
        return DCN_Mix.load(checkpoint, strict, empty_embedding=empty_embedding)
    raise NotImplementedError()


def save_ctr_checkpoint(model, checkpoint_dir: str, name: str = "target"):
    field_name = "deepfm" if isinstance(model, DeepFM) else "dcn" if isinstance(model, DCN_Mix) else None
    if field_name is None:
        raise NotImplementedError(f"Not supported for {model.__class__=}")
    field_dir = os.path.join(checkpoint_dir, field_name)
    os.makedirs(field_dir, exist_ok=True)
    torch.save(model.embedding.state_dict(), os.path.join(field_dir, f"{name}.pth"))


def install_into_reference() -> None:
    """Rebind the reference's registry / model symbols to the B200-native classes.

    Call after `import src.models` (the reference on sys.path) and before the unchanged
    script's `main()` builds its model; see INTEGRATION.md."""
    import importlib

    emb_pkg = importlib.import_module("src.models.embeddings")
    for k, v in NAME_TO_CLS.items():
        emb_pkg.NAME_TO_CLS[k] = v
    emb_pkg.VanillaEmbedding = VanillaEmbedding   # "vanilla" is special-cased by symbol (__init__.py:53-59)
    models = importlib.import_module("src.models")
    deepfm_mod = importlib.import_module("src.models.deepfm")
    dcn_mod = importlib.import_module("src.models.dcn")
    for mod in (models, deepfm_mod):
        mod.DeepFM = DeepFM
    for mod in (models, dcn_mod):
        mod.DCN_Mix = DCN_Mix
    models.get_ctr_model = get_ctr_model
    deepfm_mod.get_optimizers = get_optimizers
    tr = importlib.import_module("src.trainer.deepfm")
    tr.DeepFM = DeepFM
    # scripts import some plugin classes by symbol from their defining modules for isinstance checks
    # (scripts/deepfm/train_deepfm_optembed.py:13,45 `isinstance(model.embedding, OptEmbed)`): rebind those too
    for mod_name, names in (("src.models.embeddings.deepfm_opt_embed", ("OptEmbed", "RetrainOptEmbed", "IOptEmbed")),
                            ("src.models.embeddings.pep_embedding", ("PepEmbeeding", "RetrainPepEmbedding")),
                            ("src.models.embeddings.qr_embedding", ("QRHashingEmbedding",)),
                            ("src.models.embeddings.cerp_embedding", ("CerpEmbedding", "RetrainCerpEmbedding")),
                            ("src.models.embeddings.base", ("VanillaEmbedding",))):
        try:
            mod = importlib.import_module(mod_name)
        except ImportError:
            continue
        for n in names:
            if hasattr(mod, n):
                setattr(mod, n, globals()[n] if n in globals() else getattr(embeddings, n))
    try:    # inference-only CSR table (needs numba on the reference side); scripts import the symbol from here
        importlib.import_module("src.models.embeddings.pruned_embedding").PrunedEmbedding = PrunedEmbedding
    except ImportError:
        pass
