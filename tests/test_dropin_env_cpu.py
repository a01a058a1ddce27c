"""CPU: the environment shims under transformer-gan_b200/compat (SURVEY.md 8b "environment gaps") -- the yacs-compatible
CfgNode, and the start-up hooks that make the reference's unmodified scripts resolve the hot-path modules to this
package.  The last test needs the reference checkout and is skipped where it does not exist (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "transformer-gan_b200", "compat")
REF_MODEL = "/root/reference/model"


def _cfgnode():
    sys.path.insert(0, COMPAT)
    try:
        import importlib
        return importlib.import_module("yacs.config").CfgNode
    finally:
        sys.path.remove(COMPAT)


def test_cfgnode_merge_freeze_clone(tmp_path):
    CN = _cfgnode()
    cfg = CN()
    cfg.MODEL = CN()
    cfg.MODEL.units = 500
    cfg.MODEL.dropout = 0.1
    cfg.TRAIN = CN()
    cfg.TRAIN.lr = 0.004
    cfg.TRAIN.layers = ["0", "1"]
    f = tmp_path / "exp.yml"
    f.write_text("MODEL:\n  units: 64\nTRAIN:\n  lr: 1\n  layers: ['3']\n")
    cfg.freeze()
    cfg.merge_from_file(str(f))               # yacs merges through item assignment: allowed on a frozen node
    assert cfg.MODEL.units == 64 and cfg.TRAIN.lr == 1.0 and isinstance(cfg.TRAIN.lr, float) and cfg.TRAIN.layers == ["3"]
    with pytest.raises(AttributeError):
        cfg.MODEL.units = 3
    cfg.defrost()
    cfg.MODEL.units = 3
    c2 = cfg.clone()
    c2.MODEL.units = 7
    assert cfg.MODEL.units == 3 and c2.MODEL.units == 7
    cfg.merge_from_list(["MODEL.dropout", "0.25", "TRAIN.lr", 0.5])
    assert cfg.MODEL.dropout == 0.25 and cfg.TRAIN.lr == 0.5
    with pytest.raises(KeyError):
        cfg.merge_from_list(["MODEL.nope", 1])
    with pytest.raises(ValueError):
        cfg.merge_from_list(["MODEL.units", "abc"])
    bad = tmp_path / "bad.yml"
    bad.write_text("MODEL:\n  unknown_key: 1\n")
    with pytest.raises(KeyError):
        cfg.merge_from_file(str(bad))
    assert "units: 3" in cfg.dump()


@pytest.mark.skipif(not os.path.isdir(REF_MODEL), reason="reference checkout not present")
def test_reference_scripts_resolve_to_this_package():
    """cwd = the reference's model/ (as train.py runs), only compat/ on PYTHONPATH: the hot-path modules come from this
    package, the yacs schema of the reference's own config helper loads its experiment file, TransformerGAN builds."""
    code = textwrap.dedent("""
        import sys
        import transformer_gan, mem_transformer, discriminator
        from transformers import AdamW
        import torch
        assert AdamW is torch.optim.AdamW
        for m in (transformer_gan, mem_transformer, discriminator):
            assert "transformer-gan_b200" in m.__file__, m.__file__
        from utils.config_helper import get_default_cfg_training
        from utils.helpers import get_fixed_temperature
        from utils.bleu import BLEU  # needs the nltk stand-in
        cfg = get_default_cfg_training()
        cfg.merge_from_file("training_config/experiment_cnn.yml")
        class V:
            vec_len = 0
            def __len__(self):
                return 310
        m = transformer_gan.TransformerGAN(cfg, V())
        n = sum(p.numel() for p in m.generator.parameters())
        assert n == 13677310, n
        sd = m.generator.state_dict()
        for k in ("word_emb.emb_layers.0.weight", "layers.5.dec_attn.qkv_net.weight", "r_w_bias", "crit.out_layers.0.bias"):
            assert k in sd, k
        print("DROPIN_OK")
    """)
    env = dict(os.environ, PYTHONPATH=COMPAT)
    out = subprocess.run([sys.executable, "-c", code], cwd=REF_MODEL, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DROPIN_OK" in out.stdout, out.stderr[-3000:]
