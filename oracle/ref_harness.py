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
def injected_uniform(noise_list, alpha_list=None):
    """Make the reference's ``torch.rand(shape)`` inside sample_gumbel (mem_transformer.py:610) return the given
    tensors in order; with ``alpha_list`` the GP interpolation weights of calc_gradient_penalty
    (``torch.rand([B,1,1], device=...)``, transformer_gan.py:204) are injected the same way."""
    it = iter(noise_list)
    it_alpha = iter(alpha_list) if alpha_list is not None else None
    orig = torch.rand

    def fake_rand(*shape, **kw):
        if kw.get("device") is not None:
            if it_alpha is None:
                return orig(*shape, **kw)
            return next(it_alpha).clone().view(*shape[0]) if len(shape) == 1 else next(it_alpha).clone()
        return next(it).clone()

    torch.rand = fake_rand
    try:
        yield
    finally:
        torch.rand = orig


# ---------------------------------------------------------------------------------------------- GAN step
class TinyVocab:
    """Stand-in for BaseVocab: the GAN step only needs ``len(vocab)`` and ``vocab.vec_len`` (transformer_gan.py:127-131)."""

    def __init__(self, n, vec_len=0):
        self.n, self.vec_len = n, vec_len

    def __len__(self):
        return self.n


def make_gan_cfg(shape, tgt_len, mem_len, dis_type, dis_tgt_len, dis_mem_len, context_len, sample_chunks_mem, loss_type,
                 bert_path="", batch_chunk=1, gen_loss_factor=1.0, dis_loss_factor=1.0):
    """yacs-shaped config with every field TransformerGAN reads (utils/config_helper.py:51-147)."""
    cfg = make_cfg(shape.n_layer, shape.n_head, shape.d_model, shape.d_inner, tgt_len, mem_len,
                   same_length=shape.same_length, clamp_len=shape.clamp_len, pre_lnorm=shape.pre_lnorm)
    cfg.DISCRIMINATOR = _Node(type=dis_type, tgt_len=dis_tgt_len, mem_len=dis_mem_len, context_len=context_len,
                              sample_chunks_mem=sample_chunks_mem, truncate_backprop=False, backprop_outside=True,
                              gen_loss_factor=gen_loss_factor, dis_loss_factor=dis_loss_factor, batch_chunk=batch_chunk,
                              BERT=_Node(model_path=bert_path, loss_type=loss_type if dis_type == "bert" else "rsgan",
                                         model_type="bert_lm", random_weights=True, freeze_layers=[]),
                              CNN=_Node(embed_dim=64, hidden_dim=64, num_rep=64, init="uniform",
                                        loss_type=loss_type if dis_type == "cnn" else "rsgan"))
    cfg.PPO = _Node(dis_D_type="bert", dis_D_num_rep=1, clip_param=0.4)
    return cfg


def tiny_bert_config_dir(path, vocab_size, hidden=32, layers=2, heads=2, inter=64, max_pos=64):
    """Write a config.json for a small BERT (dropout 0) so ``BertConfig.from_pretrained(path)`` works offline."""
    import json
    os.makedirs(path, exist_ok=True)
    import txl_oracle
    cfg = dict(txl_oracle.TINY_BERT, vocab_size=vocab_size, hidden_size=hidden, num_hidden_layers=layers,
               num_attention_heads=heads, intermediate_size=inter, max_position_embeddings=max_pos)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    return path


def build_reference_gan(cfg, n_token, gen_params, dis_state=None, dtype=torch.float64):
    """Unmodified reference TransformerGAN with the generator loaded from ``gen_params`` and (optionally) the
    discriminator from ``dis_state``; returns the module in train() mode (dropout is 0 in these configs)."""
    _, tg = load_reference(with_gan=True)
    model = tg.TransformerGAN(cfg, TinyVocab(n_token)).to(dtype)
    sd = {k: v.clone().to(dtype) for k, v in gen_params.items()}
    sd["crit.out_layers.0.weight"] = sd["word_emb.emb_layers.0.weight"]
    missing, unexpected = model.generator.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    if dis_state is not None:
        model.discriminator.load_state_dict({k: v.to(dtype) for k, v in dis_state.items()}, strict=False)
    if hasattr(model.discriminator, "config"):
        # transformers >= 4.36 defaults to SDPA attention, whose double backward (needed by the reference's WGAN-GP,
        # transformer_gan.py:220-223) does not exist; the pinned 2.5.1 only had the eager math path
        model.discriminator.config._attn_implementation = "eager"
    model.train()
    return model
