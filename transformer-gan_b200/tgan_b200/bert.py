"""The BERT discriminator's encoder on the libtgan_b200 kernels (value, input gradient and forward tangent).

Reference call sites: transformer_gan.py:391-445 (``BertForSequenceClassification(inputs_embeds=...)`` on real / fake
chunks), :203-230 (``calc_gradient_penalty``), :535-585 (construction, frozen set).  The reference's arithmetic lives in
the third-party HuggingFace ``transformers`` (pinned ==2.5.1, requirements.sh:12; ``modeling_bert.py``:
BertEmbeddings, BertSelfAttention, BertSelfOutput, BertIntermediate, BertOutput); this module restates that published
algorithm on the repo's own kernels:

  * every dense layer (QKV, attention output, intermediate, output, the one-hot -> embedding projection) is
    ``tgan_gemm`` (tcgen05, bf16 operands, fp32 accumulation) with the bias / dropout / residual epilogues;
  * LayerNorm (eps 1e-12), GELU (erf), the 64-token attention tiles and the position / token-type embedding add are
    the kernels of ``csrc/bert_ops.cu`` / ``csrc/rowops.cu``.

It covers the SHIPPED trainable set (experiment_spanbert.yml: ``freeze_layers ['0'..'4']`` with pretrained embeddings
-> only pooler + classifier train, transformer_gan.py:568-585): the encoder is a fixed function, so the engine needs
  forward      h0 = Enc(x)[:, 0]                       (the pooler reads token 0 only)
  dgrad        dx = J^T dh0                            (generator update: gradient to the sampled rows; GP: grad_x D)
  jvp          dh0 = J xdot                            (GP: d/dtheta ||grad_x D|| = <J w, d a / d theta>, see GradPenalty)
and never a weight gradient or a double-backward graph.  The pooler / classifier head (592 k parameters, [B, 768]
activations) stays in torch: it is the part that trains.  When any encoder / embedding tensor requires grad the caller
falls back to the HuggingFace modules (library path, as in the reference).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import lib as L


class _Saved:
    pass


class BertEncoderEngine:
    def __init__(self, model, dtype=torch.bfloat16, seed: int = 0):
        bert = model.bert
        cfg = model.config
        self.model = model
        self.dtype = dtype
        self.seed = seed & 0x7FFFFFFFFFFFFFFF
        self.H, self.heads, self.I = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        self.dh = self.H // self.heads
        self.n_layer = cfg.num_hidden_layers
        self.eps = float(cfg.layer_norm_eps)
        self.p_hidden, self.p_att = float(cfg.hidden_dropout_prob), float(cfg.attention_probs_dropout_prob)
        self.device = bert.embeddings.word_embeddings.weight.device
        self.calls = 0
        self._packed_key = None
        self._tables = {}

    # -- eligibility ------------------------------------------------------------------------------------------
    @staticmethod
    def supported(model) -> Optional[str]:
        """None when the encoder can run on the kernels, else the reason."""
        cfg = getattr(model, "config", None)
        if cfg is None or not hasattr(model, "bert") or not hasattr(model, "classifier"):
            return "not a BertForSequenceClassification"
        H, heads = cfg.hidden_size, cfg.num_attention_heads
        if H % heads or (H // heads) % 8 or H // heads > 64 or H % 8 or cfg.intermediate_size % 8 or H > 1024:
            return "hidden / head sizes outside the kernels' range"
        if cfg.hidden_act != "gelu":
            return f"hidden_act {cfg.hidden_act}"
        if getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            return "relative position embeddings"
        return None

    def frozen(self) -> bool:
        bert = self.model.bert
        return not any(p.requires_grad for m in (bert.embeddings, bert.encoder) for p in m.parameters())

    # -- parameters -------------------------------------------------------------------------------------------
    def pack(self):
        bert = self.model.bert
        tensors = list(bert.embeddings.parameters()) + list(bert.encoder.parameters())
        key = (tuple(t.data_ptr() for t in tensors), sum(t._version for t in tensors))
        if key == self._packed_key:
            return
        dt, dev = self.dtype, self.device
        cvt = lambda t: t.detach().to(device=dev, dtype=dt).contiguous()
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        emb = bert.embeddings
        E = emb.word_embeddings.weight.detach()
        V = E.shape[0]
        self.V, self.VP = V, (V + 63) // 64 * 64
        Ep = torch.zeros(self.VP, self.H, dtype=dt, device=dev)
        Ep[:V] = E.to(dt)
        self.E = Ep                                   # [VP, H]   (rows = vocabulary)
        self.Et = Ep.t().contiguous()                 # [H, VP]   nn.Linear layout of x = onehot @ E
        self.pos_type = f32(emb.position_embeddings.weight + emb.token_type_embeddings.weight[0][None, :])
        self.emb_g, self.emb_b = f32(emb.LayerNorm.weight), f32(emb.LayerNorm.bias)
        self.layers = []
        for lyr in bert.encoder.layer:
            a, o, i, out = lyr.attention.self, lyr.attention.output, lyr.intermediate, lyr.output
            w = _Saved()
            Wqkv = torch.cat([a.query.weight, a.key.weight, a.value.weight], 0)
            w.Wqkv, w.WqkvT = cvt(Wqkv), cvt(Wqkv.t())
            w.bqkv = f32(torch.cat([a.query.bias, a.key.bias, a.value.bias], 0))
            w.Wo, w.WoT, w.bo = cvt(o.dense.weight), cvt(o.dense.weight.t()), f32(o.dense.bias)
            w.g1, w.b1 = f32(o.LayerNorm.weight), f32(o.LayerNorm.bias)
            w.Wi, w.WiT, w.bi = cvt(i.dense.weight), cvt(i.dense.weight.t()), f32(i.dense.bias)
            w.Wo2, w.Wo2T, w.bo2 = cvt(out.dense.weight), cvt(out.dense.weight.t()), f32(out.dense.bias)
            w.g2, w.b2 = f32(out.LayerNorm.weight), f32(out.LayerNorm.bias)
            self.layers.append(w)
        self._scratch_g = torch.zeros(2 * self.H, dtype=torch.float32, device=dev)  # sink of ln_bwd's dgamma / dbeta
        self._packed_key = key

    # -- helpers ----------------------------------------------------------------------------------------------
    def _buf(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=dtype or self.dtype, device=self.device)

    def _site(self, cid, local):
        return (1 << 40) + cid * 64 + local

    def _linear(self, x, Wt_or_W, out, M, N, K, bias=None, aux=None, drop_p=0.0, site=0):
        """out[M, N] = x[M, K] @ W[N, K]^T (+ bias) (dropout) (+ aux)"""
        flags = (L.EPI_BIAS if bias is not None else 0) | (L.EPI_ADD_AUX if aux is not None else 0) | \
                (L.EPI_DROPOUT if drop_p > 0 else 0)
        L.gemm(x, Wt_or_W, out, M=M, N=N, K=K, bias=bias, aux=aux, ldaux=0 if aux is None else aux.stride(0), flags=flags,
               drop_p=drop_p, seed=self.seed, site=site)

    # -- forward ----------------------------------------------------------------------------------------------
    def embed_onehot(self, soft: torch.Tensor) -> torch.Tensor:
        """soft: [B, T, V'] float rows (one-hot / relaxed) -> x = soft @ E_bert as a [B*T, H] tensor (compute dtype)"""
        self.pack()
        B, T, V = soft.shape
        R = B * T
        s = self._buf(R, self.VP)
        L.convert(soft.reshape(R, V).contiguous().float(), V, s, self.VP, R, V, self.VP)
        x = self._buf(R, self.H)
        self._linear(s, self.Et, x, R, self.H, self.VP)
        return x

    def embed_onehot_dgrad(self, dx: torch.Tensor, B: int, T: int, V: int) -> torch.Tensor:
        """dx [B*T, H] -> d soft [B, T, V] fp32"""
        R = B * T
        d = self._buf(R, self.VP, dtype=torch.float32)
        L.gemm(dx, self.E, d, M=R, N=self.VP, K=self.H)
        return d[:, :V].reshape(B, T, V)

    def forward(self, B: int, T: int, *, x: Optional[torch.Tensor] = None, ids: Optional[torch.Tensor] = None,
                training: bool, save: bool) -> _Saved:
        """x: [B*T, H] input embeddings (compute dtype) or ids: int64 [B, T].  -> ctx with .h0 (fp32 [B, H])"""
        self.pack()
        H, I, R = self.H, self.I, B * T
        self.calls += 1
        cid = self.calls
        ph = self.p_hidden if training else 0.0
        pa = self.p_att if training else 0.0
        c = _Saved()
        c.B, c.T, c.R, c.cid, c.ph, c.pa = B, T, R, cid, ph, pa
        if T > self.pos_type.shape[0]:
            raise L.TganError("sequence longer than the position table")
        z0 = self._buf(R, H, dtype=torch.float32)
        if x is not None:
            L.bert_embed_rows(z0, R, H, self.pos_type, T, x=x.contiguous())
        else:
            L.bert_embed_rows(z0, R, H, self.pos_type, T, ids=ids.reshape(-1).contiguous(), E=self.E)
        h = self._buf(R, H)
        c.z0, c.mean0, c.rstd0 = z0, self._buf(R, dtype=torch.float32), self._buf(R, dtype=torch.float32)
        L.ln_fwd_eps(z0, h, self.emb_g, self.emb_b, c.mean0, c.rstd0, R, H, self.eps)
        if ph > 0:
            L.dropout(h, h, R, H, H, H, ph, self.seed, self._site(cid, 0))
        c.layers = []
        for l, w in enumerate(self.layers):
            s = _Saved()
            s.qkv = self._buf(R, 3 * H)
            self._linear(h, w.Wqkv, s.qkv, R, 3 * H, H, bias=w.bqkv)
            s.ctx = self._buf(R, H)
            s.lse = self._buf(B * self.heads * T, dtype=torch.float32)
            L.bert_attn_fwd(s.qkv, s.ctx, s.lse, B, self.heads, T, self.dh, pa, self.seed, self._site(cid, 8 * l + 1))
            s.z1 = self._buf(R, H, dtype=torch.float32)
            self._linear(s.ctx, w.Wo, s.z1, R, H, H, bias=w.bo, aux=h, drop_p=ph, site=self._site(cid, 8 * l + 2))
            s.h1 = self._buf(R, H)
            s.mean1, s.rstd1 = self._buf(R, dtype=torch.float32), self._buf(R, dtype=torch.float32)
            L.ln_fwd_eps(s.z1, s.h1, w.g1, w.b1, s.mean1, s.rstd1, R, H, self.eps)
            s.u = self._buf(R, I)
            self._linear(s.h1, w.Wi, s.u, R, I, H, bias=w.bi)
            g = self._buf(R, I)
            L.gelu(s.u, g, R, I)
            s.z2 = self._buf(R, H, dtype=torch.float32)
            self._linear(g, w.Wo2, s.z2, R, H, I, bias=w.bo2, aux=s.h1, drop_p=ph, site=self._site(cid, 8 * l + 3))
            h = self._buf(R, H)
            s.mean2, s.rstd2 = self._buf(R, dtype=torch.float32), self._buf(R, dtype=torch.float32)
            L.ln_fwd_eps(s.z2, h, w.g2, w.b2, s.mean2, s.rstd2, R, H, self.eps)
            if save:
                c.layers.append(s)
        h0 = torch.empty(B, H, dtype=torch.float32, device=self.device)
        L.convert(h, T * H, h0, H, B, H, H)   # token 0 of every sequence: row stride T * H
        c.h0 = h0
        if not save:
            c.z0 = None
        return c

    # -- input gradient -----------------------------------------------------------------------------------------
    def dgrad(self, c: _Saved, dh0: torch.Tensor) -> torch.Tensor:
        """dh0: [B, H] gradient w.r.t. the token-0 outputs -> dx [B*T, H] (compute dtype): J^T dh0"""
        H, I, R, B, T, cid = self.H, self.I, c.R, c.B, c.T, c.cid
        dy = torch.zeros(R, H, dtype=self.dtype, device=self.device)
        L.convert(dh0.contiguous().float(), H, dy, T * H, B, H, H)
        sg = self._scratch_g
        for l in reversed(range(self.n_layer)):
            w, s = self.layers[l], c.layers[l]
            dz2 = self._buf(R, H)
            dz2d = self._buf(R, H) if c.ph > 0 else None
            L.ln_bwd(dy, s.z2, w.g2, s.mean2, s.rstd2, dz2, dz2d, sg, sg[H:], R, H, H, c.ph, self.seed,
                     self._site(cid, 8 * l + 3))
            dg = self._buf(R, I)
            self._linear(dz2d if dz2d is not None else dz2, w.Wo2T, dg, R, I, H)
            du = self._buf(R, I)
            L.gelu(s.u, du, R, I, t=dg)
            dh1 = self._buf(R, H)
            self._linear(du, w.WiT, dh1, R, H, I, aux=dz2)
            dz1 = self._buf(R, H)
            dz1d = self._buf(R, H) if c.ph > 0 else None
            L.ln_bwd(dh1, s.z1, w.g1, s.mean1, s.rstd1, dz1, dz1d, sg, sg[H:], R, H, H, c.ph, self.seed,
                     self._site(cid, 8 * l + 2))
            dctx = self._buf(R, H)
            self._linear(dz1d if dz1d is not None else dz1, w.WoT, dctx, R, H, H)
            dqkv = self._buf(R, 3 * H)
            L.bert_attn_bwd(s.qkv, dctx, s.lse, dqkv, B, self.heads, T, self.dh, c.pa, self.seed, self._site(cid, 8 * l + 1))
            dy = self._buf(R, H)
            self._linear(dqkv, w.WqkvT, dy, R, H, 3 * H, aux=dz1)
        if c.ph > 0:
            L.dropout(dy, dy, R, H, H, H, c.ph, self.seed, self._site(cid, 0))
        dx = self._buf(R, H)
        L.ln_bwd(dy, c.z0, self.emb_g, c.mean0, c.rstd0, dx, None, sg, sg[H:], R, H, H, 0.0, self.seed, 0)
        return dx

    # -- forward tangent ----------------------------------------------------------------------------------------
    def jvp(self, c: _Saved, xdot: torch.Tensor) -> torch.Tensor:
        """xdot: [B*T, H] direction in input-embedding space -> J xdot restricted to token 0: [B, H] fp32"""
        H, I, R, B, T, cid = self.H, self.I, c.R, c.B, c.T, c.cid
        hd = self._buf(R, H)
        L.ln_jvp(xdot.contiguous().float(), c.z0, self.emb_g, c.mean0, c.rstd0, hd, R, H)
        if c.ph > 0:
            L.dropout(hd, hd, R, H, H, H, c.ph, self.seed, self._site(cid, 0))
        for l, (w, s) in enumerate(zip(self.layers, c.layers)):
            qkvd = self._buf(R, 3 * H)
            self._linear(hd, w.Wqkv, qkvd, R, 3 * H, H)
            ctxd = self._buf(R, H)
            L.bert_attn_jvp(s.qkv, qkvd, s.lse, ctxd, B, self.heads, T, self.dh, c.pa, self.seed, self._site(cid, 8 * l + 1))
            z1d = self._buf(R, H, dtype=torch.float32)
            self._linear(ctxd, w.Wo, z1d, R, H, H, aux=hd, drop_p=c.ph, site=self._site(cid, 8 * l + 2))
            h1d = self._buf(R, H)
            L.ln_jvp(z1d, s.z1, w.g1, s.mean1, s.rstd1, h1d, R, H)
            ud = self._buf(R, I)
            self._linear(h1d, w.Wi, ud, R, I, H)
            gd = self._buf(R, I)
            L.gelu(s.u, gd, R, I, t=ud)
            z2d = self._buf(R, H, dtype=torch.float32)
            self._linear(gd, w.Wo2, z2d, R, H, I, aux=h1d, drop_p=c.ph, site=self._site(cid, 8 * l + 3))
            hd = self._buf(R, H)
            L.ln_jvp(z2d, s.z2, w.g2, s.mean2, s.rstd2, hd, R, H)
        out = torch.empty(B, H, dtype=torch.float32, device=self.device)
        L.convert(hd, T * H, out, H, B, H, H)
        return out


# ---------------------------------------------------------------------------------------------------------
# autograd glue
# ---------------------------------------------------------------------------------------------------------
class _EncodeFn(torch.autograd.Function):
    """x [B*T, H] -> h0 [B, H];  backward = the dgrad pass."""

    @staticmethod
    def forward(ctx, eng, x, B, T, training, holder):
        c = eng.forward(B, T, x=x.detach(), training=training, save=True)
        ctx.eng, ctx.c = eng, c
        if holder is not None:
            holder.append(c)
        # The returned tensor becomes an autograd output of this node: it must not ALSO live inside ctx (node -> ctx.c ->
        # h0 -> grad_fn = node is a reference cycle that keeps the whole graph -- and the AccumulateGrad nodes bound to the
        # stream it was built on -- alive until the garbage collector runs; a later CUDA-graph capture on another stream
        # then trips over them: "dependency created on uncaptured work in another stream").
        h0, c.h0 = c.h0, None
        return h0

    @staticmethod
    def backward(ctx, dh0):
        dx = ctx.eng.dgrad(ctx.c, dh0)
        return None, dx.to(torch.float32) if ctx.needs_input_grad[1] else None, None, None, None, None


class _EmbedOneHotFn(torch.autograd.Function):
    """soft [B, T, V'] -> x = soft @ E_bert [B*T, H];  backward: d soft = dx @ E^T"""

    @staticmethod
    def forward(ctx, eng, soft):
        ctx.eng, ctx.shape = eng, soft.shape
        return eng.embed_onehot(soft.detach()).float()

    @staticmethod
    def backward(ctx, dx):
        B, T, V = ctx.shape
        return None, ctx.eng.embed_onehot_dgrad(dx.to(ctx.eng.dtype).contiguous(), B, T, V)


class _VjpFn(torch.autograd.Function):
    """a [B, H] -> g = J^T a [B*T, H] (the dgrad pass as a FUNCTION of its seed); its adjoint is the JVP pass.  Used by
    the gradient penalty: a = d head / d h0 depends on the trainable head, J (the frozen encoder at x^) does not."""

    @staticmethod
    def forward(ctx, eng, c, a):
        ctx.eng, ctx.c = eng, c
        return eng.dgrad(c, a.detach()).float()

    @staticmethod
    def backward(ctx, w):
        return None, None, ctx.eng.jvp(ctx.c, w.contiguous())


def encode(eng: BertEncoderEngine, x: torch.Tensor, B: int, T: int, training: bool, holder=None) -> torch.Tensor:
    return _EncodeFn.apply(eng, x, B, T, training, holder)
