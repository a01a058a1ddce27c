"""Harness that imports the UNMODIFIED reference modules from /root/reference/model.

TEST INFRASTRUCTURE ONLY (see oracle/txl_oracle.py header).  Usable only where /root/reference exists
(the build container); the GPU box never has it, so nothing under ``tests -m gpu``, ``smoke()`` or
``bench.py`` may import this file.  It exists to (1) validate the restatement in txl_oracle.py against the
real reference and (2) produce the committed golden vectors (oracle/make_goldens.py).

Harness-side shims only, no edits to the reference:
  * ``yacs.config.CfgNode``   -> a tiny attribute-dict (yacs is not installed here)
  * ``transformers.AdamW``    -> torch.optim.AdamW (removed from transformers 5.x; transformer_gan.py:23-30)
  * ``torch.Tensor.cuda``     -> identity on CPU-only hosts (mem_transformer.py:610 hard-codes .cuda())
"""
from __future__ import annotations

import os
import sys
import types
from contextlib import contextmanager

import torch

REF_MODEL_DIR = "/root/reference/model"


class _Node(dict):
    """Minimal stand-in for yacs CfgNode: attribute access on a dict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def freeze(self):
        pass

    def defrost(self):
        pass


def available() -> bool:
    return os.path.isdir(REF_MODEL_DIR)


def _install_shims():
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        yacs_config = types.ModuleType("yacs.config")
        yacs_config.CfgNode = _Node
        yacs.config = yacs_config
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = yacs_config
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # type: ignore[assignment]


def load_reference(with_gan: bool = False):
    """Returns the reference module(s): mem_transformer [, transformer_gan]."""
    if not available():
        raise RuntimeError("reference tree not present: " + REF_MODEL_DIR)
    _install_shims()
    for pth in (REF_MODEL_DIR, os.path.join(REF_MODEL_DIR, "utils")):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    # The reference's top-level module names (mem_transformer, utils, ...) collide with the drop-in package's;
    # make sure the reference copies win inside this harness.
    for name in ("mem_transformer", "transformer_gan", "discriminator", "helpers", "utils",
                 "utils.proj_adaptive_softmax", "utils.helpers"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith("/root/reference"):
            del sys.modules[name]
    import mem_transformer  # noqa: E402
    if not with_gan:
        return mem_transformer
    import transformers
    if not hasattr(transformers, "AdamW"):
        transformers.AdamW = torch.optim.AdamW
    import transformer_gan  # noqa: E402
    return mem_transformer, transformer_gan


def make_cfg(n_layer, n_head, d_model, d_inner, tgt_len, mem_len, same_length=False, clamp_len=-1,
             pre_lnorm=False, dropout=0.0, dropatt=0.0):
    cfg = _Node()
    cfg.MODEL = _Node(num_layers=n_layer, num_heads=n_head, units=d_model, inner_size=d_inner, dropout=dropout,
                      attention_dropout=dropatt, tie_embedding=True, tie_proj=False, pre_lnorm=pre_lnorm,
                      same_length=same_length, clamp_len=clamp_len)
    cfg.TRAIN = _Node(tgt_length=tgt_len, mem_length=mem_len, pad_type="model", replace_start_with_pad=False,
                      append_note_status=False)
    return cfg


def build_reference_lm(shape, params, tgt_len, dtype=torch.float32):
    """Instantiate the reference MemTransformerLM with ``params`` (txl_oracle.init_params layout)."""
    mt = load_reference()
    cfg = make_cfg(shape.n_layer, shape.n_head, shape.d_model, shape.d_inner, tgt_len, shape.mem_len,
                   same_length=shape.same_length, clamp_len=shape.clamp_len, pre_lnorm=shape.pre_lnorm)
    model = mt.MemTransformerLM(cfg, shape.n_token, 0).to(dtype)
    sd = {k: v.clone().to(dtype) for k, v in params.items()}
    sd["crit.out_layers.0.weight"] = sd["word_emb.emb_layers.0.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m.endswith("inv_freq") for m in missing), missing
    model.eval()
    return model


@contextmanager
def injected_uniform(noise_list):
    """Make the reference's ``torch.rand(shape)`` inside sample_gumbel (mem_transformer.py:610) return the
    given tensors in order."""
    it = iter(noise_list)
    orig = torch.rand

    def fake_rand(*shape, **kw):
        if kw.get("device") is not None:  # calc_gradient_penalty's alpha (transformer_gan.py:204): leave alone
            return orig(*shape, **kw)
        return next(it).clone()

    torch.rand = fake_rand
    try:
        yield
    finally:
        torch.rand = orig
