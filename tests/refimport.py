"""Locate and import the UNMODIFIED reference for the tests that compare against it / run its trainer.

The reference is staged under baseline/_ref/ by `__graft_entry__.build()` (git-ignored, travels to the GPU box with
the snapshot like librsb.so); /root/reference itself is never read by a `-m gpu` test.  `lmdb` / `optuna` are absent
from the image and only touched when real dataset caches / hyper-parameter searches are opened: they are stubbed."""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RSB_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "models"))


def activate() -> str:
    """Put the reference on sys.path (its package is literally called `src`) and stub the absent dependencies."""
    if not available():
        raise RuntimeError(f"reference not staged at {REF}: run __graft_entry__.build() in the build container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
    sys.modules.setdefault("optuna", types.ModuleType("optuna"))
    try:
        from loguru import logger

        logger.remove()
    except Exception:  # noqa: BLE001
        pass
    return REF


def load_script(rel: str, name: str):
    """Import one of the reference's scripts (e.g. scripts/deepfm/train_deepfm_pep.py) as a module, unmodified."""
    activate()
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_REBOUND_MODULES = ("src.models", "src.models.embeddings", "src.models.deepfm", "src.models.dcn", "src.trainer.deepfm",
                    "src.models.embeddings.pruned_embedding", "src.models.embeddings.deepfm_opt_embed",
                    "src.models.embeddings.pep_embedding", "src.models.embeddings.qr_embedding",
                    "src.models.embeddings.cerp_embedding", "src.models.embeddings.base")


class installed:
    """`recsys_benchmark_b200.install_into_reference()` for the duration of a `with` block; the reference's own
    symbols and registry are put back afterwards so that other tests still see the unmodified reference."""

    def __enter__(self):
        activate()
        self.mods = {}
        for n in _REBOUND_MODULES:
            try:
                self.mods[n] = importlib.import_module(n)
            except ImportError:
                pass
        self.saved = {n: dict(vars(m)) for n, m in self.mods.items()}
        self.registry = dict(self.mods["src.models.embeddings"].NAME_TO_CLS)
        import recsys_benchmark_b200 as R

        R.install_into_reference()
        return self.mods

    def __exit__(self, *exc):
        for n, m in self.mods.items():
            for k, v in self.saved[n].items():
                setattr(m, k, v)
        reg = self.mods["src.models.embeddings"].NAME_TO_CLS
        reg.clear()
        reg.update(self.registry)
