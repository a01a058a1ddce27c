"""GPU: the BERT discriminator's encoder on the repo's kernels (tgan_b200/bert.py, csrc/bert_ops.cu) against the
reference's own third-party arithmetic -- HuggingFace ``BertForSequenceClassification`` (eager attention), the module
transformer_gan.py:391-445 calls -- on the same weights and inputs: value, input gradient, forward tangent, and the
WGAN-GP term with its gradient (transformer_gan.py:203-230) against autograd's double backward.
Tolerances: fp32 mode 1e-3, bf16 1e-2 relative (BASELINE.json's logits / loss tolerances; north_star)."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

CFGS = {
    "tiny": dict(hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64, T=8, B=3),
    "mid": dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, T=64, B=4),
    "wide": dict(hidden_size=192, num_hidden_layers=1, num_attention_heads=12, intermediate_size=384, T=33, B=2),
}


def make_model(name, seed=0, dropout=0.0):
    from transformers import BertConfig, BertForSequenceClassification
    c = dict(CFGS[name])
    T, B = c.pop("T"), c.pop("B")
    cfg = BertConfig(vocab_size=311, max_position_embeddings=64, type_vocab_size=2, hidden_act="gelu",
                     hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout, layer_norm_eps=1e-12, **c)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    m = BertForSequenceClassification(cfg)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # non-trivial weights everywhere (HF's init leaves biases 0, LayerNorm at identity)
        for n, p in m.named_parameters():
            if n.endswith("LayerNorm.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.12 * torch.randn(p.shape, generator=g))
    m = m.cuda().train()
    for n, p in m.named_parameters():
        p.requires_grad_(not (n.startswith("bert.embeddings") or n.startswith("bert.encoder")))
    return m, T, B


def hf_h0(m, x):
    h = m.bert.embeddings(inputs_embeds=x)
    h = m.bert.encoder(h, attention_mask=None)
    h = h[0] if isinstance(h, (tuple, list)) else h.last_hidden_state
    return h[:, 0]


def hf_logit(m, x):
    return m.classifier(m.dropout(m.bert.pooler.activation(m.bert.pooler.dense(hf_h0(m, x)))))[:, 0]


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", ["tiny", "mid", "wide"])
def test_encoder_value_dgrad_jvp_match_huggingface(name, dtype):
    from tgan_b200.bert import BertEncoderEngine
    m, T, B = make_model(name)
    H = m.config.hidden_size
    assert BertEncoderEngine.supported(m) is None
    eng = BertEncoderEngine(m, dtype, seed=1)
    assert eng.frozen()
    tol = 1e-3 if dtype == torch.float32 else 2.5e-2
    g = torch.Generator().manual_seed(3)
    x = (0.5 * torch.randn(B, T, H, generator=g)).cuda()
    dh0 = torch.randn(B, H, generator=g).cuda()
    xdot = torch.randn(B, T, H, generator=g).cuda()
    # reference: HF modules in fp32
    xr = x.clone().requires_grad_(True)
    h0_ref = hf_h0(m, xr)
    (dx_ref,) = torch.autograd.grad(h0_ref, xr, grad_outputs=dh0)
    _, jvp_ref = torch.autograd.functional.jvp(lambda t: hf_h0(m, t), x, xdot)
    c = eng.forward(B, T, x=x.reshape(B * T, H).to(dtype), training=True, save=True)
    assert rel(c.h0, h0_ref) < tol, rel(c.h0, h0_ref)
    dx = eng.dgrad(c, dh0).float().reshape(B, T, H)
    assert rel(dx, dx_ref) < tol, rel(dx, dx_ref)
    jv = eng.jvp(c, xdot.reshape(B * T, H))
    assert rel(jv, jvp_ref) < tol, rel(jv, jvp_ref)
    # ids path == embedding rows of the same ids; one-hot projection == matmul
    ids = torch.randint(0, 311, (B, T), generator=g).cuda()
    E = m.bert.embeddings.word_embeddings.weight
    h0_ids = eng.forward(B, T, ids=ids, training=True, save=False).h0
    assert rel(h0_ids, hf_h0(m, E[ids])) < tol
    soft = torch.softmax(2 * torch.randn(B, T, 311, generator=g), -1).cuda()
    xs = eng.embed_onehot(soft).float().reshape(B, T, H)
    assert rel(xs, soft @ E) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gan_losses_and_gradient_penalty_match_the_huggingface_path(dtype, tmp_path):
    """TransformerGAN with the frozen-encoder discriminator: own-kernel path vs the HuggingFace path of the same module on
    the same inputs -- D(real), D(fake) and their gradient to the fake rows, and the WGAN-GP value + its gradient w.r.t.
    the trainable head (autograd double backward through HF on one side, one JVP pass on the other)."""
    import transformer_gan as TG
    m, T, B = make_model("mid", seed=5)
    V = 310
    gan = TG.TransformerGAN.__new__(TG.TransformerGAN)
    torch.nn.Module.__init__(gan)
    ns = types.SimpleNamespace
    gan.cfg = ns(DISCRIMINATOR=ns(type="bert"))
    gan.discriminator, gan.ntokens = m, V
    gan.generator = ns(compute_dtype=dtype)
    gan.use_own_bert, gan._bert_engine = True, None
    g = torch.Generator().manual_seed(11)
    alpha = torch.rand(B, generator=g).cuda()
    gan.gp_alpha_source = lambda b: alpha
    real = torch.randint(2, V, (B, T), generator=g).cuda()
    fake = torch.softmax(3 * torch.randn(B, T, V + 1, generator=g), -1).cuda()
    real_1h = torch.zeros(B, T, V + 1, device="cuda").scatter_(-1, real[..., None], 1.0)
    tol = 2e-3 if dtype == torch.float32 else 3e-2
    eng = gan._own_bert()
    assert eng is not None
    head = [p for p in m.parameters() if p.requires_grad]
    E = m.bert.embeddings.word_embeddings.weight
    # discriminator scores + gradient to the fake rows (generator update)
    f1 = fake.clone().requires_grad_(True)
    d_fake = gan._own_logit(eng, soft=f1)
    (gf,) = torch.autograd.grad(d_fake.sum(), f1)
    f2 = fake.clone().requires_grad_(True)
    d_fake_ref = gan._bert_logit(torch.einsum("ve,bcv->bce", E, f2))
    (gf_ref,) = torch.autograd.grad(d_fake_ref.sum(), f2)
    assert rel(d_fake, d_fake_ref) < tol and rel(gf, gf_ref) < tol, (rel(d_fake, d_fake_ref), rel(gf, gf_ref))
    assert rel(gan._own_logit(eng, ids=real), gan._bert_logit(E[real])) < tol
    # gradient penalty: value and gradient w.r.t. pooler / classifier
    gp = gan._own_gradient_penalty(eng, real_1h, fake)
    grads = torch.autograd.grad(gp, head, allow_unused=True)
    gan.use_own_bert = False
    gp_ref = gan.calc_gradient_penalty(real_1h, fake)
    grads_ref = torch.autograd.grad(gp_ref, head, allow_unused=True)
    assert abs(gp.item() - gp_ref.item()) <= tol * max(1.0, abs(gp_ref.item())), (gp.item(), gp_ref.item())
    for p, a, b in zip(head, grads, grads_ref):
        if b is None:
            assert a is None or a.abs().max() == 0
            continue
        assert rel(a, b) < 4 * tol, (p.shape, rel(a, b))


def test_dropout_masks_are_consistent_between_value_dgrad_and_jvp():
    """With dropout ON the three passes must see the same masks: <J^T a, w> == <a, J w> (adjoint identity) holds only if
    the dgrad and JVP passes linearise the very function the forward pass evaluated."""
    from tgan_b200.bert import BertEncoderEngine
    m, T, B = make_model("mid", seed=7, dropout=0.1)
    H = m.config.hidden_size
    eng = BertEncoderEngine(m, torch.float32, seed=3)
    g = torch.Generator().manual_seed(5)
    x = (0.5 * torch.randn(B * T, H, generator=g)).cuda()
    a = torch.randn(B, H, generator=g).cuda()
    w = torch.randn(B * T, H, generator=g).cuda()
    c = eng.forward(B, T, x=x, training=True, save=True)
    lhs = (eng.dgrad(c, a).float() * w).sum().item()
    rhs = (a * eng.jvp(c, w)).sum().item()
    assert abs(lhs - rhs) <= 2e-3 * max(1.0, abs(rhs)), (lhs, rhs)
    # and the tangent is the derivative of the (masked) forward: finite difference along w
    eps = 1e-2
    eng.calls -= 1
    hp = eng.forward(B, T, x=x + eps * w, training=True, save=False).h0
    eng.calls -= 1
    hm = eng.forward(B, T, x=x - eps * w, training=True, save=False).h0
    fd = (hp - hm) / (2 * eps)
    assert rel(eng.jvp(c, w), fd) < 2e-2
