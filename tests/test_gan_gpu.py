"""GPU: the drop-in TransformerGAN (generator on the CUDA kernels, discriminator = HF BERT / RelGAN_D) against the
golden vectors of the UNMODIFIED reference GAN step (tests/golden/gan_*.npz) and against the oracle's sampled ids."""
import json
import os
import types

import numpy as np
import pytest
import torch

import golden_util as GU
import txl_oracle as O
from test_gan_golden import build_disc

pytestmark = pytest.mark.gpu


class _Vocab:
    vec_len = 0

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def _cfg(shape, z, bert_dir):
    ns = types.SimpleNamespace
    T = int(z["dis_tgt_len"])
    dis_type, loss_type = str(z["dis_type"]), str(z["loss_type"])
    return ns(MODEL=ns(num_layers=shape.n_layer, num_heads=shape.n_head, units=shape.d_model, inner_size=shape.d_inner,
                       dropout=0.0, attention_dropout=0.0, tie_embedding=True, tie_proj=False, pre_lnorm=False,
                       same_length=shape.same_length, clamp_len=shape.clamp_len),
              TRAIN=ns(tgt_length=T, mem_length=shape.mem_len, pad_type="model", replace_start_with_pad=False,
                       append_note_status=False),
              DISCRIMINATOR=ns(type=dis_type, tgt_len=T, mem_len=shape.mem_len, context_len=int(z["context_len"]),
                               sample_chunks_mem=int(z["chunks"]), truncate_backprop=False, backprop_outside=True,
                               gen_loss_factor=1.0, dis_loss_factor=1.0, batch_chunk=1,
                               BERT=ns(model_path=bert_dir, loss_type=loss_type if dis_type == "bert" else "rsgan",
                                       model_type="bert_lm", random_weights=True, freeze_layers=[]),
                               CNN=ns(embed_dim=64, hidden_dim=64, num_rep=64, init="uniform",
                                      loss_type=loss_type if dis_type == "cnn" else "rsgan")),
              PPO=ns(dis_D_type="bert", dis_D_num_rep=1, clip_param=0.4))


@pytest.mark.parametrize("name", ["gan_bert_tiny", "gan_cnn_tiny"])
def test_gan_step_matches_reference_golden(name, tmp_path):
    import transformer_gan as TG
    z, shape = GU.load(name)
    V, B, T, ctx, chunks = shape.n_token, int(z["B"]), int(z["dis_tgt_len"]), int(z["context_len"]), int(z["chunks"])
    bert_dir = str(tmp_path / "bert")
    os.makedirs(bert_dir, exist_ok=True)
    json.dump(dict(O.TINY_BERT, vocab_size=V + 1), open(os.path.join(bert_dir, "config.json"), "w"))
    torch.manual_seed(0)
    model = TG.TransformerGAN(_cfg(shape, z, bert_dir), _Vocab(V))
    sd = {k: v.clone() for k, v in O.init_params(shape, int(z["seed"])).items()}
    sd["crit.out_layers.0.weight"] = sd["word_emb.emb_layers.0.weight"]
    model.generator.load_state_dict(sd, strict=False)
    model.discriminator.load_state_dict(O.seeded_state(model.discriminator, int(z["seed"]) + 1), strict=False)
    if hasattr(model.discriminator, "dropout"):
        model.discriminator.dropout.p = 0.0
    model = model.cuda().train()
    model.generator.compute_dtype = torch.float32  # fp32 parity mode (1e-4)
    model.temperature = float(z["temperature"])
    data = torch.from_numpy(z["data"]).cuda()
    U = torch.from_numpy(z["U"]).cuda()
    alpha = torch.from_numpy(z["alpha"]).cuda()
    chunk_of = {"k": 0}
    model.gumbel_noise_source = lambda step, shp: U[step:step + 1]

    def alpha_src(b):
        a = alpha[chunk_of["k"] % chunks]
        chunk_of["k"] += 1
        return a
    model.gp_alpha_source = alpha_src

    # the oracle replays the same call on the CPU (fp64): source of the expected sampled ids
    disc, dparams, extra, embed, on_emb = build_disc(z, shape, torch.float64)
    Ul = [torch.from_numpy(z["U"][k:k + 1]).double() for k in range(T - ctx)]
    al = [torch.from_numpy(z["alpha"][k]).double() for k in range(chunks)]
    for mode in ("dis_loss", "gen_loss"):
        model.zero_grad(set_to_none=True)
        chunk_of["k"] = 0
        r = model(data, None, None, mode)
        torch.cuda.synchronize()
        for key in ("dis_loss", "gen_loss", "gp_loss"):
            if f"{mode}.{key}" in z.files:
                want = float(z[f"{mode}.{key}"])
                got = float(r[key])
                assert abs(got - want) <= 1e-3 * max(1.0, abs(want)), (mode, key, got, want)
        p = {k: v.double().requires_grad_(True) for k, v in O.init_params(shape, int(z["seed"])).items()}
        ro = O.gan_step(mode, torch.from_numpy(z["data"]), p, shape, disc, extra, str(z["loss_type"]),
                        float(z["temperature"]), Ul, al, T, ctx, chunks, embed=embed, disc_on_embeds=on_emb)
        # sampled ids are bit-exact under the injected noise (generated positions only)
        ids = model.last_sampled_ids.cpu()
        want_ids = torch.cat([c for c in torch.split(ro["ids"], T // chunks)], 0)
        gen_rows = [i for i in range(T) if i >= ctx]
        assert torch.equal(ids, want_ids[gen_rows]), (mode, (ids != want_ids[gen_rows]).sum().item())
        owner = model.discriminator if mode == "dis_loss" else model.generator
        named = dict(owner.named_parameters())
        checked = 0
        for k in z.files:
            pre = f"{mode}.grad."
            if not k.startswith(pre) or k[len(pre):] == "crit.out_layers.0.weight":
                continue
            nm = k[len(pre):]
            want = torch.from_numpy(z[k]).double()
            g = named[nm].grad
            got = g.detach().cpu().double() if g is not None else torch.zeros_like(want)
            err = (got - want).norm().item()
            assert err <= 2e-2 * want.norm().item() + 2e-6, (mode, nm, err, want.norm().item())
            checked += 1
        assert checked >= 5
        if mode == "dis_loss":
            assert all(prm.grad is None for prm in model.generator.parameters())


def _build_gan(name, tmp_path, dtype):
    import transformer_gan as TG
    z, shape = GU.load(name)
    V = shape.n_token
    bert_dir = str(tmp_path / "bert")
    os.makedirs(bert_dir, exist_ok=True)
    json.dump(dict(O.TINY_BERT, vocab_size=V + 1), open(os.path.join(bert_dir, "config.json"), "w"))
    torch.manual_seed(0)
    model = TG.TransformerGAN(_cfg(shape, z, bert_dir), _Vocab(V))
    sd = {k: v.clone() for k, v in O.init_params(shape, int(z["seed"])).items()}
    sd["crit.out_layers.0.weight"] = sd["word_emb.emb_layers.0.weight"]
    model.generator.load_state_dict(sd, strict=False)
    model.discriminator.load_state_dict(O.seeded_state(model.discriminator, int(z["seed"]) + 1), strict=False)
    if hasattr(model.discriminator, "dropout"):
        model.discriminator.dropout.p = 0.0
    model = model.cuda().train()
    model.generator.compute_dtype = dtype
    model.temperature = float(z["temperature"])
    return model, z, shape


@pytest.mark.parametrize("name", ["gan_bert_tiny", "gan_cnn_tiny"])
def test_gan_step_bf16_eager_and_graphed_match_reference(name, tmp_path):
    """The configuration bench.py times: generator in bf16, projected-K/V cache, the whole adversarial phase replayed as
    one CUDA graph (TransformerGAN.use_cuda_graphs).  Losses within 1e-2 (bf16 tolerance of BASELINE.json) of the
    UNMODIFIED reference's goldens; sampled ids equal to the oracle's along every sequence up to the first step whose
    top-2 (logit + g) margin is inside the tolerance; the graph replay reproduces the eager bf16 call."""
    model, z, shape = _build_gan(name, tmp_path, torch.bfloat16)
    V, B, T, ctx, chunks = shape.n_token, int(z["B"]), int(z["dis_tgt_len"]), int(z["context_len"]), int(z["chunks"])
    data = torch.from_numpy(z["data"]).cuda()
    U = torch.from_numpy(z["U"]).cuda()
    alpha = torch.from_numpy(z["alpha"]).cuda()
    chunk_of = {"k": 0}
    model.gumbel_noise_source = lambda step, shp: U[step:step + 1]

    def alpha_src(b):
        a = alpha[chunk_of["k"] % chunks]
        chunk_of["k"] += 1
        return a
    model.gp_alpha_source = alpha_src
    model.sources_graph_safe = True  # views of static device tensors
    disc, dparams, extra, embed, on_emb = build_disc(z, shape, torch.float64)
    Ul = [torch.from_numpy(z["U"][k:k + 1]).double() for k in range(T - ctx)]
    al = [torch.from_numpy(z["alpha"][k]).double() for k in range(chunks)]
    gen_rows = [i for i in range(T) if i >= ctx]

    def run(mode):
        model.zero_grad(set_to_none=False)
        chunk_of["k"] = 0
        r = model(data, None, None, mode)
        torch.cuda.synchronize()
        owner = model.discriminator if mode == "dis_loss" else model.generator
        grads = {k: p.grad.detach().clone() for k, p in owner.named_parameters() if p.grad is not None}
        return {k: float(v) for k, v in r.items() if v is not None and k != "mems"}, model.last_sampled_ids.clone(), grads

    for mode in ("dis_loss", "gen_loss"):
        model.use_cuda_graphs = False
        for prm in model.parameters():
            if prm.grad is None and prm.requires_grad:
                prm.grad = torch.zeros_like(prm)
        r_eager, ids_eager, g_eager = run(mode)
        p = {k: v.double().requires_grad_(True) for k, v in O.init_params(shape, int(z["seed"])).items()}
        ro = O.gan_step(mode, torch.from_numpy(z["data"]), p, shape, disc, extra, str(z["loss_type"]),
                        float(z["temperature"]), Ul, al, T, ctx, chunks, embed=embed, disc_on_embeds=on_emb)
        want_ids = torch.cat([c for c in torch.split(ro["ids"], T // chunks)], 0)[gen_rows]
        margins = ro["margins"]  # [T - ctx, B]
        ids = ids_eager.cpu()
        all_on_prefix = True
        for b in range(B):
            for t in range(T - ctx):
                if margins[t, b] < 2e-2:  # inside the bf16 tolerance: this and later steps of the column may differ
                    all_on_prefix = False
                    break
                assert ids[t, b] == want_ids[t, b], (mode, t, b, float(margins[t, b]))
        if all_on_prefix:  # same samples -> same discriminator inputs: the losses are comparable at the bf16 tolerance
            for key in ("dis_loss", "gen_loss", "gp_loss"):
                if f"{mode}.{key}" in z.files:
                    want, got = float(z[f"{mode}.{key}"]), r_eager[key]
                    assert abs(got - want) <= 1e-2 * max(1.0, abs(want)), (mode, key, got, want)
        # graphed: warm call (eager), capturing call, replay -- each must reproduce the eager bf16 numbers
        model.use_cuda_graphs = True
        for rep in range(3):
            r_g, ids_g, g_g = run(mode)
            assert torch.equal(ids_g, ids_eager), (mode, rep)
            for key, want in r_eager.items():
                assert abs(r_g[key] - want) <= 2e-3 * max(1.0, abs(want)), (mode, rep, key, r_g[key], want)
            for k, want in g_eager.items():
                err = (g_g[k] - want).norm().item()
                assert err <= 2e-2 * want.norm().item() + 1e-6, (mode, rep, k, err, want.norm().item())
        assert any(k[0] == mode for k in model._gan_graphs), "the adversarial phase was not captured"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gan_ppo_variant_matches_reference_golden(dtype, tmp_path):
    """The PPO variant (loss type 'ppo-gp', density-ratio classifier dis_D = RelGAN_D): 'classifier_loss', 'gen_loss'
    with and without update_D0, 'dis_loss' -- against the golden of the UNMODIFIED reference (transformer_gan.py:133-153,
    :184-201, :350-388).  fp32 mode at 1e-3; bf16 at 1e-2 when the sampled ids coincide."""
    import transformer_gan as TG
    from test_gan_golden import PPO_CALLS
    z, shape = GU.load("gan_ppo_tiny")
    V, B, T, ctx, chunks = shape.n_token, int(z["B"]), int(z["dis_tgt_len"]), int(z["context_len"]), int(z["chunks"])
    bert_dir = str(tmp_path / "bert")
    os.makedirs(bert_dir, exist_ok=True)
    json.dump(dict(O.TINY_BERT, vocab_size=V + 1), open(os.path.join(bert_dir, "config.json"), "w"))
    cfg = _cfg(shape, z, bert_dir)
    cfg.PPO.dis_D_type, cfg.PPO.clip_param = "cnn", float(z["clip"])
    torch.manual_seed(0)
    model = TG.TransformerGAN(cfg, _Vocab(V))
    sd = {k: v.clone() for k, v in O.init_params(shape, int(z["seed"])).items()}
    sd["crit.out_layers.0.weight"] = sd["word_emb.emb_layers.0.weight"]
    model.generator.load_state_dict(sd, strict=False)
    model.discriminator.load_state_dict(O.seeded_state(model.discriminator, int(z["seed"]) + 1), strict=False)
    model.dis_D.load_state_dict(O.seeded_state(model.dis_D, int(z["seed"]) + 3), strict=False)
    model.dis_D.dropout.p = 0.0
    model = model.cuda().train()
    model.generator.compute_dtype = dtype
    model.temperature = float(z["temperature"])
    model.use_cuda_graphs = True  # must be ignored for the PPO variants (host-side P0 / update_D0 state)
    data = torch.from_numpy(z["data"]).cuda()
    U = torch.from_numpy(z["U"]).cuda()
    alpha = torch.from_numpy(z["alpha"]).cuda()
    chunk_of = {"k": 0}
    model.gumbel_noise_source = lambda step, shp: U[step:step + 1]

    def alpha_src(b):
        a = alpha[chunk_of["k"] % chunks]
        chunk_of["k"] += 1
        return a
    model.gp_alpha_source = alpha_src
    fp32 = dtype == torch.float32
    tol = 1e-3 if fp32 else 1e-2
    ids_ref = None
    for tag, mode, upd in PPO_CALLS:
        model.zero_grad(set_to_none=True)
        chunk_of["k"] = 0
        r = model(data, None, None, mode, update_D0=upd)
        torch.cuda.synchronize()
        if ids_ref is None:
            ids_ref = model.last_sampled_ids.clone()  # same noise, same weights: every call samples the same ids
        same_ids = torch.equal(model.last_sampled_ids, ids_ref)
        assert same_ids or not fp32
        if mode == "classifier_loss":
            assert all(r[k] is None for k in ("mle", "gen_loss", "dis_loss"))
        if not same_ids:
            continue
        want_P0 = torch.from_numpy(z[f"{tag}.P0"]).cuda()
        assert torch.allclose(model.P0.float(), want_P0, rtol=10 * tol, atol=1e-4), (tag, model.P0, want_P0)
        for key in ("dis_loss", "gen_loss", "gp_loss"):
            if f"{tag}.{key}" in z.files:
                want, got = float(z[f"{tag}.{key}"]), float(r[key])
                assert abs(got - want) <= tol * max(1.0, abs(want)), (tag, key, got, want)
        owner = {"classifier_loss": model.dis_D, "dis_loss": model.discriminator}.get(mode, model.generator)
        named = dict(owner.named_parameters())
        checked = 0
        for k in z.files:
            pre = f"{tag}.grad."
            if not k.startswith(pre) or k[len(pre):] == "crit.out_layers.0.weight":
                continue
            want = torch.from_numpy(z[k]).double()
            g = named[k[len(pre):]].grad
            got = g.detach().cpu().double() if g is not None else torch.zeros_like(want)
            err = (got - want).norm().item()
            # (key-bias gradients are mathematically zero -- softmax shift invariance: absolute floor; TF32 / bf16
            # rounding leaves ~1e-4 there)
            assert err <= (2e-2 if fp32 else 0.15) * want.norm().item() + (2e-6 if fp32 else 5e-4), (tag, k, err, want.norm().item())
            checked += 1
        assert checked >= 5
    assert not model._gan_graphs
