"""Drop-in check against the unmodified reference staged under baseline/_ref (by __graft_entry__.build(); it travels
to the GPU box with the snapshot): after install_into_reference() the reference's own registry / factories hand out
the B200-native classes, with the reference's yaml configs."""
import os

import pytest

from tests import refimport

REF = refimport.REF
pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference not staged under baseline/_ref")


def test_registry_is_rebound_and_reference_yaml_builds_our_model():
    import yaml

    import recsys_benchmark_b200 as R

    with refimport.installed() as mods:
        ref_models, ref_emb = mods["src.models"], mods["src.models.embeddings"]
        assert ref_emb.NAME_TO_CLS["qr"] is R.QRHashingEmbedding
        assert ref_emb.VanillaEmbedding is R.VanillaEmbedding
        assert ref_models.DeepFM is R.DeepFM and ref_models.DCN_Mix is R.DCN_Mix
        assert mods["src.models.embeddings.deepfm_opt_embed"].OptEmbed is R.OptEmbed
        # the reference's own factory (unchanged code) now returns our plugin classes
        emb = ref_emb.get_embedding({"name": "vanilla"}, [3, 4], 8)
        assert isinstance(emb, R.VanillaEmbedding)
        emb = ref_emb.get_embedding({"name": "qr", "divider": 2}, [3, 4], 8)
        assert isinstance(emb, R.QRHashingEmbedding)
        assert ref_emb.NAME_TO_CLS["dhe"] is R.DHEmbedding and ref_emb.NAME_TO_CLS["cerp"] is R.CerpEmbedding
        assert mods["src.models.embeddings.pruned_embedding"].PrunedEmbedding is R.PrunedEmbedding
        # the DHE yaml (k = 1024, 4 x 1536 encoder) builds our plugin; no 4 GB code cache is materialised
        with open(os.path.join(REF, "configs", "deepfm", "dhe_config-50.yaml")) as fh:
            dhe_cfg = yaml.safe_load(fh)["model"]
        dhe_cfg["embedding_config"].pop("cache_path", None)
        dhe_model = ref_models.get_ctr_model([5, 6, 7], dict(dhe_cfg))
        assert isinstance(dhe_model.embedding, R.DHEmbedding) and dhe_model.embedding._inp_size == 1024
        for cfg_name in ["base_config.yaml", "qr_80.yaml", "base_config_sparse.yaml"]:
            with open(os.path.join(REF, "configs", "deepfm", cfg_name)) as fh:
                cfg = yaml.safe_load(fh)
            model = ref_models.get_ctr_model([5, 6, 7], dict(cfg["model"]))
            assert isinstance(model, R.DeepFM)
            opts = ref_models.deepfm.get_optimizers(model, cfg)
            assert len(opts) == (2 if cfg.get("sparse") else 1)
    # ... and the reference's own symbols are back afterwards
    import src.models as ref_models

    assert ref_models.DeepFM.__module__.startswith("src.")
