"""CPU oracle for the Transformer-XL generator / GAN-step hot path.

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (torch on the host, fp32 or fp64) of the
algorithm in the reference's ``model/mem_transformer.py``, ``model/utils/proj_adaptive_softmax.py``
and the sampling part of ``model/transformer_gan.py``.  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as
the checker or the timed CPU baseline, never as the product path.  The product path
(``transformer-gan_b200/``) never imports it and fails loudly when the CUDA library is missing.

Parity status: **pinned**.  The reference's own tests hold no numerical fixture for this path
(SURVEY.md section 4), so the pins are golden vectors produced by running the unmodified reference
modules in the build container (``oracle/make_goldens.py`` -> ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks this restatement against every one of them.

The restatement is deliberately *not* a transcription: the reference builds the relative shift with a
pad/reshape trick and the mask as a materialised bool tensor; here both are closed-form index
arithmetic, which is also what the CUDA kernels implement.  Gradients come from
autograd over this restatement (fp64 capable).

All ``file:line`` citations are relative to ``/root/reference/model``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class TxlShape:
    """Hyper-parameters of MemTransformerLM (mem_transformer.py:351-367)."""
    n_layer: int
    n_head: int
    d_model: int
    d_inner: int
    n_token: int
    mem_len: int
    same_length: bool = False
    clamp_len: int = -1
    pre_lnorm: bool = False

    @property
    def d_head(self) -> int:
        return self.d_model // self.n_head  # mem_transformer.py:354


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def param_names(shape: TxlShape) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict names / shapes of the generator (SURVEY.md section 5 'Checkpoint / resume')."""
    D, N, dh, DI, V = shape.d_model, shape.n_head, shape.d_head, shape.d_inner, shape.n_token
    out = [("r_w_bias", (N, dh)), ("r_r_bias", (N, dh)), ("word_emb.emb_layers.0.weight", (V, D))]
    for i in range(shape.n_layer):
        p = f"layers.{i}."
        out += [
            (p + "dec_attn.qkv_net.weight", (3 * N * dh, D)),
            (p + "dec_attn.o_net.weight", (D, N * dh)),
            (p + "dec_attn.layer_norm.weight", (D,)),
            (p + "dec_attn.layer_norm.bias", (D,)),
            (p + "dec_attn.r_net.weight", (N * dh, D)),
            (p + "pos_ff.CoreNet.0.weight", (DI, D)),
            (p + "pos_ff.CoreNet.0.bias", (DI,)),
            (p + "pos_ff.CoreNet.3.weight", (D, DI)),
            (p + "pos_ff.CoreNet.3.bias", (D,)),
            (p + "pos_ff.layer_norm.weight", (D,)),
            (p + "pos_ff.layer_norm.bias", (D,)),
        ]
    out += [("crit.out_layers.0.bias", (V,))]  # crit.out_layers.0.weight is tied to the embedding (:411-413)
    return out


def init_params(shape: TxlShape, seed: int, std: float = 0.02, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic parameters (CPU generator).  Not the reference's init distribution on
    purpose: std 0.02 and non-zero biases / LN offsets make every term observable in parity tests."""
    g = torch.Generator().manual_seed(seed)
    params = {}
    for name, shp in param_names(shape):
        if name.endswith("layer_norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("bias") and "r_" not in name:
            t = 0.05 * torch.randn(shp, generator=g)
        elif name in ("r_w_bias", "r_r_bias"):
            t = 0.2 * torch.randn(shp, generator=g)
        elif "emb_layers" in name:
            t = 0.05 * torch.randn(shp, generator=g)
        else:
            t = std * 2.5 * torch.randn(shp, generator=g)
        params[name] = t.to(dtype)
    return params


# ----------------------------------------------------------------------------------------------
# A1  positional embedding                                     mem_transformer.py:13-23, 550-555
# ----------------------------------------------------------------------------------------------
def positional_embedding(klen: int, d_model: int, clamp_len: int = -1, dtype=torch.float32) -> torch.Tensor:
    """pos_emb[p] = [sin(d_p f), cos(d_p f)], d_p = klen-1-p, f_t = 10000^(-2t/D).  Returns [klen, D]."""
    # the reference builds inv_freq once, in float32, as a registered buffer (mem_transformer.py:13-14); later
    # .to(dtype)/.double() only widens those rounded values -- reproduce that rounding exactly.
    inv_freq = (1 / (10000 ** (torch.arange(0.0, d_model, 2.0) / d_model))).to(dtype)
    pos_seq = torch.arange(klen - 1, -1, -1.0, dtype=dtype)
    if clamp_len > 0:
        pos_seq = pos_seq.clamp(max=clamp_len)
    ang = pos_seq[:, None] * inv_freq[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=-1)


# ----------------------------------------------------------------------------------------------
# A3  attention mask (True = masked), closed form             mem_transformer.py:495-547
# ----------------------------------------------------------------------------------------------
def attn_mask(qlen: int, mlen: int, mem_len: int, same_length: bool,
              reset_mems: Optional[torch.Tensor], bsz: int) -> torch.Tensor:
    """[B, Q, K] bool.  masked iff j > i+M  or  (same_length and j <= i - msl)  or  (reset[b] and j < M)."""
    klen = qlen + mlen
    i = torch.arange(qlen)[:, None]
    j = torch.arange(klen)[None, :]
    m = j > i + mlen                                           # triu(ones, 1+mlen)           :525-527
    if same_length:
        mask_len = klen - mem_len                              #                               :499-503
        msl = qlen - mask_len if mask_len > 0 else qlen
        m = m | (j <= i - msl)                                 # tril(ones, -mask_shift_len)   :520
    m = m[None].repeat(bsz, 1, 1)
    if reset_mems is not None and mlen > 0:
        m[reset_mems.bool().cpu(), :, :mlen] = True                  #                               :529
    return m


# ----------------------------------------------------------------------------------------------
# A4  relative-position attention (post-LN / pre-LN)           mem_transformer.py:162-257
# ----------------------------------------------------------------------------------------------
def rel_shift_gather(bd_raw: torch.Tensor, qlen: int) -> torch.Tensor:
    """_rel_shift (mem_transformer.py:133-147) as index arithmetic: out[..., i, j] = x[..., i, j+Q-1-i].
    Entries with j+Q-1-i >= K (always masked: j > i+M) are set to 0 instead of the reference's wrap-around
    garbage."""
    klen = bd_raw.shape[-1]
    i = torch.arange(qlen, device=bd_raw.device)[:, None]
    j = torch.arange(klen, device=bd_raw.device)[None, :]
    src = j + qlen - 1 - i
    valid = src < klen
    src = src.clamp(max=klen - 1)
    out = torch.gather(bd_raw, -1, src.expand(bd_raw.shape[:-2] + src.shape))
    return out * valid.to(out.dtype)


def rel_attn(w: torch.Tensor, mem: Optional[torch.Tensor], pos_emb: torch.Tensor, p: Dict[str, torch.Tensor],
             prefix: str, r_w_bias: torch.Tensor, r_r_bias: torch.Tensor, mask: torch.Tensor,
             shape: TxlShape) -> torch.Tensor:
    """w [Q,B,D], mem [M,B,D] or None, pos_emb [K,D] -> [Q,B,D]."""
    Q, B, D = w.shape
    N, dh = shape.n_head, shape.d_head
    Wqkv = p[prefix + "dec_attn.qkv_net.weight"]
    Wq, Wk, Wv = Wqkv[: N * dh], Wqkv[N * dh: 2 * N * dh], Wqkv[2 * N * dh:]      # chunk(3)   :173
    ln_w, ln_b = p[prefix + "dec_attn.layer_norm.weight"], p[prefix + "dec_attn.layer_norm.bias"]
    x = w if mem is None or mem.numel() == 0 else torch.cat([mem, w], 0)            #            :166
    xin, win = x, w
    if shape.pre_lnorm:                                                               #            :167-168
        xin = F.layer_norm(x, (D,), ln_w, ln_b)
        win = xin[-Q:]
    K = x.shape[0]
    q = (win @ Wq.t()).view(Q, B, N, dh)               # Q-projection of memory rows is discarded   :174
    k = (xin @ Wk.t()).view(K, B, N, dh)
    v = (xin @ Wv.t()).view(K, B, N, dh)
    r = (pos_emb @ p[prefix + "dec_attn.r_net.weight"].t()).view(K, N, dh)           #            :171
    ac = torch.einsum("ibnd,jbnd->bnij", q + r_w_bias, k)                            #            :201-204
    bd = rel_shift_gather(torch.einsum("ibnd,jnd->bnij", q + r_r_bias, r), Q)        #            :206-210
    s = (ac + bd) * (1.0 / math.sqrt(dh))                                            #            :213-214
    s = s.masked_fill(mask[:, None], float("-inf"))                                  #            :225
    prob = torch.softmax(s, dim=-1)                                                  #            :228
    a = torch.einsum("bnij,jbnd->ibnd", prob, v).reshape(Q, B, N * dh)               #            :239-244
    out = a @ p[prefix + "dec_attn.o_net.weight"].t()                                #            :247
    if shape.pre_lnorm:
        return w + out                                                               #            :252
    return F.layer_norm(w + out, (D,), ln_w, ln_b)                                   #            :255


# ----------------------------------------------------------------------------------------------
# A5  position-wise FFN                                        mem_transformer.py:46-60
# ----------------------------------------------------------------------------------------------
def pos_ff(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, shape: TxlShape) -> torch.Tensor:
    D = x.shape[-1]
    ln_w, ln_b = p[prefix + "pos_ff.layer_norm.weight"], p[prefix + "pos_ff.layer_norm.bias"]
    inp = F.layer_norm(x, (D,), ln_w, ln_b) if shape.pre_lnorm else x
    h = torch.relu(inp @ p[prefix + "pos_ff.CoreNet.0.weight"].t() + p[prefix + "pos_ff.CoreNet.0.bias"])
    y = h @ p[prefix + "pos_ff.CoreNet.3.weight"].t() + p[prefix + "pos_ff.CoreNet.3.bias"]
    if shape.pre_lnorm:
        return y + x
    return F.layer_norm(x + y, (D,), ln_w, ln_b)


# ----------------------------------------------------------------------------------------------
# A2 + A8 + A7  embedding, layer stack, memory update          mem_transformer.py:319-341, 484-576, 445-482
# ----------------------------------------------------------------------------------------------
def embed(inp: torch.Tensor, E: torch.Tensor, d_model: int) -> torch.Tensor:
    """index path (2-D int64 [Q,B]) or soft path (3-D float [Q,B,V]); both scaled by sqrt(D)."""
    e = E[inp] if inp.dim() == 2 else inp.to(E.dtype) @ E
    return e * math.sqrt(d_model)


def update_mems(hids: List[torch.Tensor], mems: Optional[torch.Tensor], mem_len: int) -> torch.Tensor:
    """new_mems = last mem_len rows of [mems; hids] per slab, detached (mem_transformer.py:461-475)."""
    stacked = torch.stack(hids).detach()
    cat = stacked if mems is None or mems.numel() == 0 else torch.cat([mems.detach(), stacked], 1)
    end = cat.shape[1]
    return cat[:, max(0, end - mem_len): end]


def core_forward(inp: torch.Tensor, reset_mems: Optional[torch.Tensor], mems: Optional[torch.Tensor],
                 p: Dict[str, torch.Tensor], shape: TxlShape) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """MemTransformerLM._forward with dropout 0.  Returns (core_out [Q,B,D], new_mems [L+1,M',B,D])."""
    E = p["word_emb.emb_layers.0.weight"]
    x = embed(inp, E, shape.d_model)
    Q, B = x.shape[0], x.shape[1]
    M = 0 if mems is None or mems.numel() == 0 else mems.shape[1]
    K = Q + M
    # .to(x.device): the eager-GPU bar of bench.py runs this restatement unchanged on CUDA tensors
    mask = attn_mask(Q, M, shape.mem_len, shape.same_length, reset_mems, B).to(x.device)
    pe = positional_embedding(K, shape.d_model, shape.clamp_len, dtype=x.dtype).to(x.device)
    hids = [x]
    for l in range(shape.n_layer):
        pre = f"layers.{l}."
        mem_l = None if M == 0 else mems[l]
        x = rel_attn(x, mem_l, pe, p, pre, p["r_w_bias"], p["r_r_bias"], mask, shape)
        x = pos_ff(x, p, pre, shape)
        hids.append(x)
    new_mems = update_mems(hids, mems, shape.mem_len) if shape.mem_len > 0 else None
    return x, new_mems


# ----------------------------------------------------------------------------------------------
# A9  logits + NLL                                             proj_adaptive_softmax.py:50-84
# ----------------------------------------------------------------------------------------------
def logits_of(hidden: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    return hidden @ p["word_emb.emb_layers.0.weight"].t() + p["crit.out_layers.0.bias"]


def mle_forward(data: torch.Tensor, target: torch.Tensor, reset_mems: Optional[torch.Tensor],
                mems: Optional[torch.Tensor], p: Dict[str, torch.Tensor], shape: TxlShape):
    """MemTransformerLM.forward (mem_transformer.py:653-670): returns (nll [Q,B], new_mems)."""
    hid, new_mems = core_forward(data, reset_mems, mems, p, shape)
    T = target.shape[0]
    lg = logits_of(hid[-T:].reshape(-1, shape.d_model), p)
    nll = torch.logsumexp(lg, -1) - lg.gather(1, target.reshape(-1, 1)).squeeze(1)
    return nll.view(T, -1), new_mems


def generate_logits(data: torch.Tensor, mems, p, shape: TxlShape):
    """forward_generate (mem_transformer.py:578-600): returns (logits [T,B,V], new_mems)."""
    hid, new_mems = core_forward(data, None, mems, p, shape)
    return logits_of(hid, p), new_mems


# ----------------------------------------------------------------------------------------------
# A10  Gumbel-softmax straight-through                         mem_transformer.py:609-628
# ----------------------------------------------------------------------------------------------
def gumbel_noise(U: torch.Tensor, eps: float = 1e-20) -> torch.Tensor:
    return -torch.log(-torch.log(U + eps) + eps)


def gumbel_st(logits: torch.Tensor, U: torch.Tensor, temperature: float):
    """Returns (st_out, y_soft, ids).  st_out = (onehot(argmax y) - y).detach() + y."""
    y = torch.softmax((logits + gumbel_noise(U)) / temperature, dim=-1)
    ids = y.argmax(dim=-1)
    hard = F.one_hot(ids, y.shape[-1]).to(y.dtype)
    return (hard - y).detach() + y, y, ids


def generate_gumbel(data, temperature, mems, p, shape: TxlShape, U: torch.Tensor):
    """forward_generate_gumbel with injected uniform noise U [T,B,V]."""
    lg, new_mems = generate_logits(data, mems, p, shape)
    st, y, ids = gumbel_st(lg, U, temperature)
    return st, new_mems, lg, ids


# ----------------------------------------------------------------------------------------------
# A11 (generator side)  the sampling loop of TransformerGAN.forward      transformer_gan.py:273-349
# ----------------------------------------------------------------------------------------------
def sample_fake_chunks(data: torch.Tensor, p, shape_gen: TxlShape, temperature: float, noise: List[torch.Tensor],
                       tgt_len: int, context_len: int, sample_chunks_mem: int, truncate_backprop: bool = False,
                       margins: Optional[List[torch.Tensor]] = None, hard_inputs: bool = False):
    """Replays the generator side of one 'gen_loss'/'dis_loss' call.  ``shape_gen.mem_len`` must be
    DISCRIMINATOR.mem_len (the call runs under reset_length(1, mem_len), transformer_gan.py:251).
    ``noise[k]`` is the uniform tensor [1,B,V] consumed by the k-th forward_generate_gumbel call.
    Yields (chunk_start, fake_chunk [len,B,V]) with the autograd graph of each chunk intact (memory is
    detached between steps by update_mems, and the chunk's first generated token restarts from a hard id)."""
    V = shape_gen.n_token
    seq: List[torch.Tensor] = []
    mems = None
    with torch.no_grad():
        if context_len > 1:                                                        # :281-290
            _, mems = generate_logits(data[: context_len - 1], None, p, shape_gen)
    sample_len = tgt_len // sample_chunks_mem
    k = 0
    out = []
    for cs in range(0, tgt_len, sample_len):
        ce = min(cs + sample_len, tgt_len)
        for ind in range(cs, ce):
            if ind < context_len:
                seq.append(F.one_hot(data[ind], V).to(p["r_w_bias"].dtype))             # :304-306
                continue
            if truncate_backprop or ind == cs or hard_inputs:   # :311 ('classifier' feeds hard ids)
                inp = seq[-1].argmax(-1)[None, :].detach()                            # :315
            else:
                inp = seq[-1][None]                                                   # :319
            st, mems, lg, _ = generate_gumbel(inp, temperature, mems, p, shape_gen, noise[k])
            if margins is not None:  # top-2 margin of (logit + g) per sequence: where a lower-precision run may differ
                top2 = (lg.detach()[0] + gumbel_noise(noise[k][0].to(lg.dtype))).topk(2, dim=-1).values
                margins.append(top2[:, 0] - top2[:, 1])
            k += 1
            seq.append(st[0])
        if len(seq) == sample_len + 1:                                                # :339-340
            seq = seq[1:]
        out.append((cs, torch.stack(seq, 0)))
        mems = mems.detach()                                                          # :507
        seq = [seq[-1]]
    return out


# ----------------------------------------------------------------------------------------------
# A11-A15  the adversarial part of TransformerGAN.forward            transformer_gan.py:232-533
# ----------------------------------------------------------------------------------------------
def adv_losses(d_real: torch.Tensor, d_fake: torch.Tensor, loss_type: str):
    """(g_loss, d_loss) for the objectives the shipped configs use (utils/helpers.py:117-129)."""
    if "wgan" in loss_type:                                             # experiment_spanbert.yml: 'wgan-gp'
        return -d_fake.mean(), -d_real.mean() + d_fake.mean()
    if "rsgan" in loss_type:                                            # experiment_cnn.yml default
        bce = F.binary_cross_entropy_with_logits
        return bce(d_fake - d_real, torch.ones_like(d_fake)), bce(d_real - d_fake, torch.ones_like(d_real))
    if "ppo" in loss_type:                                              # 'ppo' / 'ppo-gp', utils/helpers.py:131-136
        W = (d_fake.shape[0] * F.softmax(d_fake.detach(), dim=0)).detach()
        return -d_fake.mean(), (W * d_fake - d_real).mean()
    raise NotImplementedError(loss_type)


def gradient_penalty(disc_on_onehot, real_1h: torch.Tensor, fake: torch.Tensor, alpha: torch.Tensor, lam: float = 10.0,
                     embed=None, disc_on_embeds=None):
    """10 * mean_b (||d D / d x||_2 - 1)^2 with 1e-12 inside the sqrt (transformer_gan.py:203-230), where
    x^ = alpha real + (1 - alpha) fake on one-hot rows [B, T, V'] and alpha: [B].
    CNN discriminator: the gradient is taken w.r.t. x^ itself.  BERT discriminator: the reference re-binds
    ``interpolates`` to the EMBEDDED rows x^ E (:211-216) before differentiating, so the norm is over [T, hidden]:
    pass ``embed`` (x^ -> x^ E) and ``disc_on_embeds``."""
    B = real_1h.shape[0]
    a = alpha.to(real_1h.dtype).view(B, 1, 1)
    x = (a * real_1h + (1 - a) * fake.detach()).detach().requires_grad_(True)  # Variable(requires_grad=True), :208-209
    if embed is not None:
        x = embed(x)
        d = disc_on_embeds(x)
    else:
        d = disc_on_onehot(x)
    (g,) = torch.autograd.grad(d, x, grad_outputs=torch.ones_like(d), create_graph=True, retain_graph=True)
    slopes = torch.sqrt(g.reshape(B, -1).pow(2).sum(1) + 1e-12)
    return ((slopes - 1.0) ** 2).mean() * lam


def gan_step(mode: str, data: torch.Tensor, p, shape_gen: TxlShape, disc_on_onehot, extra_col: int, loss_type: str,
             temperature: float, noise: List[torch.Tensor], alphas: List[torch.Tensor], tgt_len: int, context_len: int,
             sample_chunks_mem: int, batch_chunk: int = 1, gen_loss_factor: float = 1.0, dis_loss_factor: float = 1.0,
             embed=None, disc_on_embeds=None, ppo: Optional[dict] = None):
    """One ``"dis_loss"`` or ``"gen_loss"`` call of TransformerGAN.forward with ``backprop_outside`` (the shipped
    setting): samples ``sample_chunks_mem`` chunks with the generator, scores real / fake with ``disc_on_onehot``
    ([B, T, V + extra_col] -> logits [B or B*rep]), back-propagates each chunk's scaled loss immediately
    (:487-502) and returns the detached sums exactly as the reference does (:515-531).
    ``extra_col`` = 1 for the BERT discriminator (its vocabulary has one more id, :396-399), 0 for the CNN one.

    PPO variants (loss_type 'ppo' / 'ppo-gp'; :184-201, :350-388): ``ppo`` = {"dis_D": callable on sequence-major
    [T, B] ids or [T, B, V] rows -> logits (``dis_D_forward``), "P0": tensor or None (the module's ``self.P0``
    state; updated in place under the key), "update_D0": bool, "clip": PPO.clip_param}.  ``mode`` may then also be
    ``"classifier_loss"``: the density-ratio classifier's BCE update on real / sampled chunks (hard-id inputs, :311)."""
    V = shape_gen.n_token
    share = batch_chunk * sample_chunks_mem
    margins: List[torch.Tensor] = []
    chunks = sample_fake_chunks(data, p, shape_gen, temperature, noise, tgt_len, context_len, sample_chunks_mem,
                                margins=margins, hard_inputs="classifier" in mode)
    g_sum = d_sum = gp_sum = 0.0
    ids = []

    def d0_ratio(fake_chunk):                                                              # :351-354 / :377-380
        with torch.no_grad():
            D0 = torch.sigmoid(ppo["dis_D"](fake_chunk))
            return (1.0 - D0) / torch.clamp(D0, min=1e-7)

    for k, (cs, fake) in enumerate(chunks):
        T = fake.shape[0]
        if mode == "dis_loss":
            fake = fake.detach()
        ids.append(fake.detach().argmax(-1))
        if "classifier" in mode:                                                           # :350-372
            if ppo["P0"] is None:
                ppo["P0"] = d0_ratio(fake)
            n = ppo["P0"].shape[0]
            err = F.binary_cross_entropy(torch.sigmoid(ppo["dis_D"](data[cs:cs + T])), fake.new_ones(n)) + \
                F.binary_cross_entropy(torch.sigmoid(ppo["dis_D"](fake.detach())), fake.new_zeros(n))
            (err / share).backward()
            continue
        ratio = None
        if mode == "gen_loss" and ppo is not None and "ppo" in loss_type:                  # :375-388
            if ppo["P0"] is None or ppo.get("update_D0", False):
                ppo["P0"] = d0_ratio(fake)
            D1 = torch.sigmoid(ppo["dis_D"](fake))
            ratio = (1.0 - D1) / torch.clamp(D1 * ppo["P0"], min=1e-7)
            ratio_clipped = torch.clamp(ratio, 1.0 - ppo["clip"], 1.0 + ppo["clip"])
        real = data[cs:cs + T].transpose(0, 1)                                            # [B, T]
        fake_bt = fake.transpose(0, 1)
        if extra_col:
            fake_bt = torch.cat([fake_bt, fake_bt.new_zeros(*fake_bt.shape[:-1], extra_col)], -1)
        real_1h = F.one_hot(real, V + extra_col).to(fake_bt.dtype)
        d_real, d_fake = disc_on_onehot(real_1h), disc_on_onehot(fake_bt)
        if ratio is not None:                                                              # :419-424 / :456-461
            surr1, surr2 = ratio * d_fake, ratio_clipped * d_fake
            d_fake = torch.where(d_fake > 0, torch.min(surr1, surr2), torch.max(surr1, surr2))
        g_loss, d_loss = adv_losses(d_real, d_fake, loss_type)
        g_sum += g_loss.detach()
        d_sum += d_loss.detach()
        if mode == "dis_loss":
            (d_loss * dis_loss_factor / share).backward()
            if "gp" in loss_type:
                gp = gradient_penalty(disc_on_onehot, real_1h, fake_bt, alphas[k], embed=embed, disc_on_embeds=disc_on_embeds)
                gp_sum += gp.detach()
                (gp * dis_loss_factor / share).backward()
        else:
            (g_loss * gen_loss_factor / share).backward()
    out = {"ids": torch.cat(ids, 0), "margins": torch.stack(margins, 0) if margins else None}
    if "classifier" in mode:
        return out
    if mode == "dis_loss":
        out["dis_loss"] = dis_loss_factor * d_sum / sample_chunks_mem
        if "gp" in loss_type:
            out["gp_loss"] = dis_loss_factor * gp_sum / sample_chunks_mem
    else:
        out["gen_loss"] = gen_loss_factor * g_sum / sample_chunks_mem
    return out


def seeded_state(module, seed: int, std: float = 0.15) -> Dict[str, torch.Tensor]:
    """Deterministic, 'responsive' discriminator weights for the GAN fixtures: every floating tensor of
    ``module.state_dict()`` is drawn from its own generator keyed by (seed, tensor name), so the reference module in
    the build container and the drop-in module on the GPU box get identical values without shipping them."""
    import zlib
    sd = {}
    for k, v in module.state_dict().items():
        if not v.is_floating_point():
            continue
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(k.encode()))
        if "LayerNorm.weight" in k:
            sd[k] = 1.0 + 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = std * torch.randn(v.shape, generator=g)
    return sd


# small BERT discriminator used by the GAN fixtures (dropout 0 so the reference run is deterministic)
TINY_BERT = dict(architectures=["BertForMaskedLM"], model_type="bert", hidden_size=32, num_hidden_layers=2,
                 num_attention_heads=2, intermediate_size=64, hidden_act="gelu", hidden_dropout_prob=0.0,
                 attention_probs_dropout_prob=0.0, max_position_embeddings=64, type_vocab_size=2, initializer_range=0.02,
                 layer_norm_eps=1e-12, pad_token_id=0)


def relgan_d_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_rep: int = 64) -> torch.Tensor:
    """Functional restatement of RelGAN_D.forward in eval mode (transformer_gan.py:90-119): x [B, T, V] ->
    logits [B * num_rep].  ``sd`` holds the module's state_dict tensors."""
    emb = (x @ sd["embeddings.weight"].t()).unsqueeze(1)                              # [B, 1, T, embed_dim]   :96-98
    single = sd["embeddings.weight"].shape[0] // num_rep
    pooled = []
    for i in range(4):                                                                 # filter sizes 2..5      :100-105
        c = F.relu(F.conv2d(emb, sd[f"convs.{i}.weight"], sd[f"convs.{i}.bias"], stride=(1, single)))
        pooled.append(c.amax(dim=2))                                                   # max over time
    feat = torch.cat(pooled, 1).permute(0, 2, 1).reshape(-1, 1200)                     # :106-109
    hw = feat @ sd["highway.weight"].t() + sd["highway.bias"]                         # :110-114
    gate = torch.sigmoid(hw)
    feat = gate * F.relu(hw) + (1.0 - gate) * feat
    out = feat @ sd["feature2out.weight"].t() + sd["feature2out.bias"]               # dropout is identity in eval
    return (out @ sd["out2logits.weight"].t() + sd["out2logits.bias"]).squeeze(1)     # :116-117


# ----------------------------------------------------------------------------------------------
# f1  generation post-processing                                generate.py:228-304
# ----------------------------------------------------------------------------------------------
def generation_probs(logits: torch.Tensor, *, temperature: float, technique: str, topk: Optional[int] = 32, p: float = 0.0,
                     exclude_bos: bool = True, suppress_empty: bool = False, empty_bar_token: int = -1) -> torch.Tensor:
    """The distribution generate.py samples the next token from, for ONE sequence: logits [V] -> probs [V].
    Restates generate.py:231-296 step by step (slice out the excluded entries, temperature, softmax, pad them back as
    zeros, then the topk / nucleus filter with renormalisation)."""
    lg = logits.clone()
    if exclude_bos:                                                                  # :232-233
        lg = lg[1:]
    if suppress_empty:                                                               # :235-247
        e = empty_bar_token - 1 if exclude_bos else empty_bar_token
        lg = torch.cat([lg[:e], lg[e + 1:]], 0)
    if temperature == 0:                                                             # :250-253
        probs = torch.zeros_like(lg)
        probs[lg.argmax()] = 1.0
    else:
        probs = F.softmax(lg / temperature, dim=-1)                                  # :255-259
    if exclude_bos:                                                                  # :261-262
        probs = F.pad(probs, [1, 0])
    if suppress_empty:                                                               # :264-266
        probs = torch.cat([probs[:empty_bar_token], F.pad(probs[empty_bar_token:], [1, 0])], 0)
    if technique in ("topk", "random"):                                              # :268-275
        if technique == "topk" and topk is not None:
            _, top_idx = torch.topk(probs, topk)
            mask = torch.zeros_like(probs)
            mask[top_idx] = 1.0
            probs = probs * mask
            probs = probs / probs.sum()
    elif technique == "nucleus":                                                     # :277-296
        if p > 0:
            sorted_probs, sorted_indices = torch.sort(probs, descending=True)
            cumulative = torch.cumsum(sorted_probs, dim=0)
            remove = cumulative >= p
            remove[1:] = remove[:-1].clone()
            remove[0] = False
            to_remove = remove.scatter(dim=0, index=sorted_indices, src=remove)
            probs = probs.clone()
            probs[to_remove] = 0
            probs = probs / probs.sum()
    else:
        raise NotImplementedError(technique)
    return probs


def categorical_from_uniform(probs: torch.Tensor, u: float):
    """torch.multinomial(probs, 1) (generate.py:302) with its RNG replaced by an injected uniform: the inverse CDF.
    Returns (token, margin) where margin = distance of u * total to the nearest CDF step (ids of a lower-precision
    implementation may differ only when that margin is inside its rounding)."""
    c = torch.cumsum(probs.double(), 0)
    target = u * c[-1]
    nz = probs > 0
    idx = int(torch.nonzero((c > target) & nz)[0]) if bool(((c > target) & nz).any()) else int(torch.nonzero(nz)[-1])
    steps = torch.cat([torch.zeros(1, dtype=torch.float64), c])
    margin = float((steps - target).abs().min())
    return idx, margin


# ----------------------------------------------------------------------------------------------
# f2  LAMB                                                                        lamb.py:57-118
# ----------------------------------------------------------------------------------------------
def lamb_step(params: List[torch.Tensor], grads: List[torch.Tensor], exp_avg: List[torch.Tensor],
              exp_avg_sq: List[torch.Tensor], lr: float, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.0,
              adam: bool = False) -> List[float]:
    """One ``Lamb.step`` in place (lamb.py:64-116; "paper v3", no bias correction); returns the trust ratios."""
    ratios = []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])                                    # :87
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])                             # :89
        weight_norm = p.norm(p=2).clamp(0, 10)                                          # :97
        adam_step = m / v.sqrt().add(eps)                                               # :99
        if weight_decay != 0:
            adam_step = adam_step + weight_decay * p                                    # :100-101
        adam_norm = adam_step.norm(p=2)                                                 # :103
        trust = 1.0 if (weight_norm == 0 or adam_norm == 0) else float(weight_norm / (adam_norm + eps))  # :105-108
        ratios.append(trust)
        p.add_(adam_step, alpha=-lr * (1.0 if adam else trust))                         # :113-116
    return ratios
