"""CPU: the oracle's restatement of the GAN step (oracle/txl_oracle.gan_step) against golden vectors produced by the
UNMODIFIED reference TransformerGAN (oracle/make_goldens.py::run_gan_case, fp32, injected Gumbel noise / GP alphas)."""
import os

import numpy as np
import pytest
import torch

import txl_oracle as O
from golden_util import GOLD


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    s = z["shape"]
    shape = O.TxlShape(n_layer=int(s[0]), n_head=int(s[1]), d_model=int(s[2]), d_inner=int(s[3]), n_token=int(s[4]),
                       mem_len=int(s[5]), same_length=bool(s[6]), clamp_len=int(s[7]), pre_lnorm=bool(s[8]))
    return z, shape


def make_tiny_bert(vocab, dtype):
    from transformers import BertConfig, BertForSequenceClassification
    cfg = BertConfig(**dict(O.TINY_BERT, vocab_size=vocab))
    cfg._attn_implementation = "eager"
    return BertForSequenceClassification(cfg).to(dtype)


def build_disc(z, shape, dtype):
    """-> (callable on [B, T, V'] one-hot-ish rows, named parameter dict, extra vocab column)."""
    seed = int(z["seed"]) + 1
    if str(z["dis_type"]) == "bert":
        m = make_tiny_bert(shape.n_token + 1, dtype)
        m.load_state_dict({k: v.to(dtype) for k, v in O.seeded_state(m, seed).items()}, strict=False)
        m.train()
        E = m.bert.embeddings.word_embeddings.weight
        on_emb = lambda e: m(inputs_embeds=e)[0][:, 0]
        return (lambda x: on_emb(x @ E)), dict(m.named_parameters()), 1, (lambda x: x @ E), on_emb

    class Shell(torch.nn.Module):  # only to enumerate RelGAN_D's state_dict names / shapes for seeded_state
        def __init__(self, V):
            super().__init__()
            self.embeddings = torch.nn.Linear(V, 64, bias=False)
            self.convs = torch.nn.ModuleList([torch.nn.Conv2d(1, 300, (f, 1), stride=(1, 1)) for f in (2, 3, 4, 5)])
            self.highway = torch.nn.Linear(1200, 1200)
            self.feature2out = torch.nn.Linear(1200, 100)
            self.out2logits = torch.nn.Linear(100, 1)

    sd = {k: v.to(dtype).requires_grad_(True) for k, v in O.seeded_state(Shell(shape.n_token), seed).items()}
    return (lambda x: O.relgan_d_forward(sd, x)), sd, 0, None, None


@pytest.mark.parametrize("name", ["gan_bert_tiny", "gan_cnn_tiny"])
def test_oracle_gan_step_matches_reference(name):
    z, shape = _load(name)
    dtype = torch.float64
    B, T, ctx, chunks = int(z["B"]), int(z["dis_tgt_len"]), int(z["context_len"]), int(z["chunks"])
    data = torch.from_numpy(z["data"])
    U = [torch.from_numpy(z["U"][k:k + 1]).to(dtype) for k in range(T - ctx)]
    alphas = [torch.from_numpy(z["alpha"][k]).to(dtype) for k in range(chunks)]
    disc, dparams, extra, embed, on_emb = build_disc(z, shape, dtype)
    for mode in ("dis_loss", "gen_loss"):
        p = {k: v.requires_grad_(True) for k, v in O.init_params(shape, int(z["seed"]), dtype=torch.float32).items()}
        p = {k: v.detach().to(dtype).requires_grad_(True) for k, v in p.items()}
        for t in dparams.values():
            t.grad = None
        r = O.gan_step(mode, data, p, shape, disc, extra, str(z["loss_type"]), float(z["temperature"]), U, alphas, T, ctx,
                       chunks, embed=embed, disc_on_embeds=on_emb)
        for key in ("dis_loss", "gen_loss", "gp_loss"):
            if f"{mode}.{key}" in z:
                want = float(z[f"{mode}.{key}"])
                assert abs(float(r[key]) - want) <= 2e-4 * max(1.0, abs(want)), (mode, key, float(r[key]), want)
        owner = dparams if mode == "dis_loss" else p
        checked = 0
        for k in z.files:
            pre = f"{mode}.grad."
            if not k.startswith(pre):
                continue
            nm = k[len(pre):]
            if nm == "crit.out_layers.0.weight":
                continue
            want = torch.from_numpy(z[k]).double()
            got = owner[nm].grad if owner[nm].grad is not None else torch.zeros_like(want)
            if nm == "word_emb.emb_layers.0.weight":  # tied: the reference accumulates the output-layer gradient here
                pass
            # (key-bias gradients are mathematically zero -- softmax shift invariance -- hence the absolute floor)
            err = (got - want).norm().item()
            assert err <= 5e-3 * want.norm().item() + 2e-6, (mode, nm, err, want.norm().item())
            checked += 1
        assert checked >= 5
        if mode == "dis_loss":
            assert all(t.grad is None for t in p.values())


class _RelGanShell(torch.nn.Module):
    """RelGAN_D's state_dict names / shapes (transformer_gan.py:44-88) for seeded_state; ``single`` = embed_dim / num_rep."""

    def __init__(self, V, single=1):
        super().__init__()
        self.embeddings = torch.nn.Linear(V, 64, bias=False)
        self.convs = torch.nn.ModuleList([torch.nn.Conv2d(1, 300, (f, single), stride=(1, single)) for f in (2, 3, 4, 5)])
        self.highway = torch.nn.Linear(1200, 1200)
        self.feature2out = torch.nn.Linear(1200, 100)
        self.out2logits = torch.nn.Linear(100, 1)


def build_ppo_parts(z, shape, dtype):
    """Discriminator (tiny BERT, 'ppo-gp') + density-ratio classifier dis_D (RelGAN_D with ONE representation) of the
    gan_ppo_tiny fixture, weights regenerated from the fixture's seeds."""
    disc, dparams, extra, embed, on_emb = build_disc(z, shape, dtype)
    V = shape.n_token
    sd = {k: v.to(dtype).requires_grad_(True) for k, v in O.seeded_state(_RelGanShell(V, 64), int(z["seed"]) + 3).items()}

    def dis_D(chunk):  # dis_D_forward, transformer_gan.py:184-201 (cnn): sequence-major ids or rows
        x = chunk.transpose(0, 1)
        if x.dim() == 2:
            x = torch.nn.functional.one_hot(x, V).to(dtype)
        return O.relgan_d_forward(sd, x, num_rep=1)
    return disc, dparams, extra, embed, on_emb, dis_D, sd


PPO_CALLS = [("classifier_loss", "classifier_loss", False), ("gen_loss_d0", "gen_loss", True),
             ("gen_loss", "gen_loss", False), ("dis_loss", "dis_loss", False)]


def test_oracle_ppo_variant_matches_reference():
    """'classifier_loss' -> 'gen_loss' (update_D0) -> 'gen_loss' -> 'dis_loss' on the unmodified reference's PPO path."""
    z, shape = _load("gan_ppo_tiny")
    dtype = torch.float64
    T, ctx, chunks = int(z["dis_tgt_len"]), int(z["context_len"]), int(z["chunks"])
    data = torch.from_numpy(z["data"])
    U = [torch.from_numpy(z["U"][k:k + 1]).to(dtype) for k in range(T - ctx)]
    alphas = [torch.from_numpy(z["alpha"][k]).to(dtype) for k in range(chunks)]
    disc, dparams, extra, embed, on_emb, dis_D, dsd = build_ppo_parts(z, shape, dtype)
    ppo = {"dis_D": dis_D, "P0": None, "clip": float(z["clip"])}
    for tag, mode, upd in PPO_CALLS:
        p = {k: v.detach().to(dtype).requires_grad_(True)
             for k, v in O.init_params(shape, int(z["seed"]), dtype=torch.float32).items()}
        for t in list(dparams.values()) + list(dsd.values()):
            t.grad = None
        ppo["update_D0"] = upd
        r = O.gan_step(mode, data, p, shape, disc, extra, "ppo-gp", float(z["temperature"]), U, alphas, T, ctx, chunks,
                       embed=embed, disc_on_embeds=on_emb, ppo=ppo)
        assert torch.allclose(ppo["P0"], torch.from_numpy(z[f"{tag}.P0"]).to(dtype), rtol=2e-4, atol=1e-6), tag
        for key in ("dis_loss", "gen_loss", "gp_loss"):
            if f"{tag}.{key}" in z:
                want = float(z[f"{tag}.{key}"])
                assert abs(float(r[key]) - want) <= 2e-4 * max(1.0, abs(want)), (tag, key, float(r[key]), want)
        owner = {"classifier_loss": dsd, "dis_loss": dparams}.get(mode, p)
        checked = 0
        for k in z.files:
            pre = f"{tag}.grad."
            if not k.startswith(pre) or k[len(pre):] == "crit.out_layers.0.weight":
                continue
            nm = k[len(pre):]
            want = torch.from_numpy(z[k]).double()
            got = owner[nm].grad if owner[nm].grad is not None else torch.zeros_like(want)
            err = (got - want).norm().item()
            assert err <= 5e-3 * want.norm().item() + 2e-6, (tag, nm, err, want.norm().item())
            checked += 1
        assert checked >= 5, (tag, checked)
