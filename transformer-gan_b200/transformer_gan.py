"""Drop-in replacement of the reference's ``model/transformer_gan.py`` (GAN training step) on the libtgan_b200 kernels.

``TransformerGAN(cfg, vocab)`` keeps the reference's constructor, attributes (``generator``, ``discriminator``,
``temperature``, ``cfg``, ``vocab``, ``vec_len``, ``ntokens``) and ``forward(data, target, reset_mems, train_loss,
mems=None, status_vec=None, update_D0=False)`` -> ``{"mle", "gen_loss", "dis_loss", "mems"[, "gp_loss"]}``
(transformer_gan.py:123-533), including its habit of running ``backward()`` inside ``forward`` and returning detached
sums (:487-502, :515-531).

What runs where
  * generator: every one of the ``tgt_len - context_len`` single-token sampling steps (:299-334) is the CUDA engine
    (``MemTransformerLM.forward_generate_gumbel`` -> tgan_* kernels, fused Gumbel-softmax straight-through sampler with
    device-side noise); the soft one-hot chain that carries the gradient through time (:308-320) is the engine's
    soft-input path (``tgan_gemm`` embedding + ``__dinput__`` gradient).
  * discriminator (BERT): with the SHIPPED trainable set (embeddings + every encoder layer frozen, pooler +
    classifier train; experiment_spanbert.yml ``freeze_layers``) the 5 x 768 encoder runs on the repo's own kernels
    (``tgan_b200.bert.BertEncoderEngine``: tcgen05 GEMMs, own LayerNorm / GELU / attention kernels) in three passes --
    value, input gradient, forward tangent -- and the WGAN-GP term (:203-230) is computed WITHOUT a double-backward
    graph: d/dtheta ||grad_x D|| only involves the encoder through one Jacobian-vector product.  Only the 592 k-parameter
    head stays in torch.  Any other trainable set falls back to HuggingFace ``BertForSequenceClassification``
    (third-party in the reference as well, transformer_gan.py:23-30 / requirements.sh:12) with eager attention so that
    autograd's double backward works.  The CNN ``RelGAN_D`` (discriminator.py) runs its dense layers on ``tgan_gemm`` in bf16
    mode (time convolutions / pooling: torch).
Differences from the reference, all numerically neutral:
  * in ``"dis_loss"`` mode the sampling loop runs under ``no_grad`` (the reference builds the 123-step graph and then
    detaches it, :346-347);
  * Gumbel noise comes from the device-side counter RNG unless ``gumbel_noise_source`` is set (parity tests inject the
    reference's uniform draws); the GP interpolation weights likewise honour ``gp_alpha_source``.
The PPO variants (loss type 'ppo' / 'ppo-gp': the density-ratio classifier ``dis_D``, the ``"classifier_loss"`` phase, the
clipped-ratio generator target, :133-153, :184-201, :350-388) keep their host-side state (``P0``, ``update_D0``) and
are therefore launched from the host rather than replayed as a graph; their arithmetic is torch on [B]-sized tensors
plus the same generator / discriminator kernels.
"""
import torch
import torch.nn as nn

from transformers import BertConfig, BertForMaskedLM, BertForSequenceClassification

from discriminator import RelGAN_D
from mem_transformer import MemTransformerLM
from utils.helpers import get_losses


def gen_is_bf16(gen) -> bool:
    return getattr(gen, "compute_dtype", torch.float32) == torch.bfloat16


class TransformerGAN(nn.Module):
    def __init__(self, cfg, vocab):
        super().__init__()
        self.ntokens = len(vocab)
        self.generator = MemTransformerLM(cfg, self.ntokens, vocab.vec_len)
        dcfg = cfg.DISCRIMINATOR
        self.cfg = cfg
        self.ppo = "ppo" in dcfg.CNN.loss_type or "ppo" in dcfg.BERT.loss_type
        if self.ppo:  # density-ratio classifier of the PPO variants (transformer_gan.py:133-153), registered before
            self.dis_D = self._create_dis_D(cfg)  # the discriminator like the reference does (state_dict order)
            self.P0 = None
        if dcfg.type == "bert":
            self.discriminator = self.create_bert_model(dcfg.BERT.model_path, dcfg.BERT.loss_type, dcfg.BERT.model_type,
                                                        dcfg.BERT.random_weights)
            self.discriminator.unfreeze_idx = self.calculate_unfreeze_idx(cfg)
        elif dcfg.type == "cnn":
            self.discriminator = RelGAN_D(dcfg.CNN.embed_dim, dcfg.tgt_len, dcfg.CNN.num_rep, self.ntokens, 1, cfg=cfg)
        else:
            self.discriminator = None
        self.cfg = cfg
        self.temperature = 1
        self.vocab = vocab
        self.vec_len = vocab.vec_len
        # test hooks: callables returning the uniform noise of the k-th sampling step ([1, B, V]) / the GP alphas ([B])
        self.gumbel_noise_source = None
        self.gp_alpha_source = None
        # the injected sources return views of static device tensors (no host work): safe to capture in a CUDA graph
        self.sources_graph_safe = False
        self.last_sampled_ids = None
        # replay the adversarial phase of forward() as one CUDA graph per (phase, batch shape); see _gan_phase_graphed
        self.use_cuda_graphs = False
        self._gan_graphs, self._gan_warm = {}, set()
        self.disc_tf32 = True  # TF32 tensor-core GEMMs for the discriminator when the generator computes in bf16
        self.use_own_bert = True   # frozen-encoder BERT discriminator on the repo's kernels (else: HuggingFace modules)
        self._bert_engine = None
        self.use_own_cnn_gemm = True  # RelGAN_D's dense layers on tgan_gemm in bf16 mode (see discriminator.py)

    # ------------------------------------------------------------------------------------------------ discriminator
    def create_bert_model(self, model_name_or_path, loss_type, model_type=None, random_weights=False):
        """transformer_gan.py:535-566.  Eager attention: the fused SDPA kernels have no double backward (WGAN-GP)."""
        config = BertConfig.from_pretrained(model_name_or_path, cache_dir=None)
        config._attn_implementation = "eager"
        if model_type == "bert_lm":
            model = BertForSequenceClassification(config=config)
            if not random_weights:
                lm = BertForMaskedLM.from_pretrained(model_name_or_path, config=config, cache_dir=None)
                # transformer_gan.py:549-553 does `model.bert = lm.bert`.  With transformers 2.5.1 (the reference's pin)
                # BertForMaskedLM's BertModel carries a pooler; current releases build it with add_pooling_layer=False,
                # so the wholesale assignment would drop the pooler the classification head needs: take the pretrained
                # embeddings + encoder (+ the pooler when the checkpoint model has one) and keep the rest.
                model.bert.embeddings = lm.bert.embeddings
                model.bert.encoder = lm.bert.encoder
                if getattr(lm.bert, "pooler", None) is not None:
                    model.bert.pooler = lm.bert.pooler
        else:
            if random_weights:
                raise NotImplementedError
            model = BertForSequenceClassification.from_pretrained(model_name_or_path, config=config, cache_dir=None)
        return model.bert if loss_type == "mmd" else model

    def _create_dis_D(self, cfg):
        """transformer_gan.py:133-149.  (For dis_D_type 'bert' the reference computes the unfreeze list from
        ``self.discriminator`` before that attribute exists, :140; here it is computed on dis_D itself.)"""
        dcfg = cfg.DISCRIMINATOR
        if cfg.PPO.dis_D_type == "bert":
            dis_D = self.create_bert_model(dcfg.BERT.model_path, dcfg.BERT.loss_type, dcfg.BERT.model_type)
            dis_D.unfreeze_idx = self.calculate_unfreeze_idx(cfg, dis_D)
            return dis_D
        if cfg.PPO.dis_D_type == "cnn":
            return RelGAN_D(dcfg.CNN.embed_dim, dcfg.tgt_len, cfg.PPO.dis_D_num_rep, self.ntokens, 1, cfg=cfg)
        raise NotImplementedError(cfg.PPO.dis_D_type)

    def dis_D_forward(self, data):
        """Logits of the density-ratio classifier on a sequence-major chunk: [T, B] ids or [T, B, V] one-hot-ish rows
        (transformer_gan.py:184-201).  The BERT variant looks hard ids up (argmax: no gradient to the generator), the
        CNN variant consumes the rows themselves (the PPO ratio is differentiable w.r.t. the samples)."""
        data = data.transpose(0, 1)
        if self.cfg.PPO.dis_D_type == "bert":
            if data.dim() == 3:
                data = data.argmax(dim=-1)
            return self._bert_logit(self.dis_D.bert.embeddings.word_embeddings.weight[data], self.dis_D)
        if data.dim() == 2:
            data = self._one_hot(data)
        return self.dis_D(data)

    def _ppo_P0(self, fake):
        with torch.no_grad():
            D0 = torch.sigmoid(self.dis_D_forward(fake))
            return (1.0 - D0) / torch.clamp(D0, min=1e-7)

    def calculate_unfreeze_idx(self, cfg, module=None):
        """Indices (in ``named_parameters`` order) of the discriminator tensors that train (transformer_gan.py:568-585)."""
        frozen_layers = cfg.DISCRIMINATOR.BERT.freeze_layers
        idx, layers = [], []
        for i, (name, _) in enumerate((module if module is not None else self.discriminator).named_parameters()):
            in_layer = name.startswith("bert.encoder.layer")
            if in_layer:
                layers.append(name.split(".")[3])
            frozen = (name.startswith("bert.embeddings") and not cfg.DISCRIMINATOR.BERT.random_weights) or \
                     (in_layer and name.split(".")[3] in frozen_layers)
            if not frozen:
                idx.append(i)
        assert len(layers) >= len(frozen_layers)
        return idx

    def _bert_embedding_matrix(self):
        return self.discriminator.bert.embeddings.word_embeddings.weight

    def _bert_logit(self, inputs_embeds, m=None):
        """Logit column 0 of BertForSequenceClassification(inputs_embeds=...) (transformer_gan.py:403-416), calling the
        sub-modules directly: with no padding the all-ones attention mask HuggingFace builds adds exactly 0.0 to every
        score, and its construction (a host scalar copied to the device) cannot be captured in a CUDA graph."""
        m = self.discriminator if m is None else m
        bert = m.bert
        h = bert.embeddings(inputs_embeds=inputs_embeds)
        h = bert.encoder(h, attention_mask=None)
        h = h[0] if isinstance(h, (tuple, list)) else h.last_hidden_state
        return m.classifier(m.dropout(bert.pooler(h)))[:, 0]

    # ------------------------------------------------------------------------------------------------ own BERT path
    def _own_bert(self):
        """The kernel-side encoder when it applies: BERT discriminator on CUDA, sizes the kernels cover, and nothing
        inside embeddings / encoder trainable (the shipped configuration).  None -> HuggingFace path."""
        if not self.use_own_bert or self.cfg.DISCRIMINATOR.type != "bert":
            return None
        m = self.discriminator
        from tgan_b200.bert import BertEncoderEngine
        if BertEncoderEngine.supported(m) is not None:
            return None
        dev = m.classifier.weight.device
        if dev.type != "cuda":
            return None
        dt = getattr(self.generator, "compute_dtype", torch.bfloat16)
        eng = self._bert_engine
        if eng is None or eng.device != dev or eng.dtype != dt or eng.model is not m:
            eng = self._bert_engine = BertEncoderEngine(m, dt, seed=torch.initial_seed())
        return eng if eng.frozen() else None

    def _bert_head(self, h0):
        """pooler (dense + tanh on token 0) -> dropout -> classifier, logit column 0 (modeling_bert.py BertPooler /
        BertForSequenceClassification.forward).  The trainable part: stays in torch."""
        m = self.discriminator
        pooled = m.bert.pooler.activation(m.bert.pooler.dense(h0))
        return m.classifier(m.dropout(pooled))[:, 0]

    def _own_logit(self, eng, *, ids=None, soft=None):
        """D(real ids [B, T]) or D(fake rows [B, T, V+1]) with the encoder on the kernels."""
        from tgan_b200 import bert as TB
        training = self.discriminator.training
        if ids is not None:
            B, T = ids.shape
            return self._bert_head(eng.forward(B, T, ids=ids, training=training, save=False).h0)
        B, T, _ = soft.shape
        if not (soft.requires_grad and torch.is_grad_enabled()):
            x = eng.embed_onehot(soft)
            return self._bert_head(eng.forward(B, T, x=x, training=training, save=False).h0)
        x = TB._EmbedOneHotFn.apply(eng, soft)
        return self._bert_head(TB.encode(eng, x, B, T, training))

    def _own_gradient_penalty(self, eng, real_1h, fake_bt, LAMBDA=10):
        """WGAN-GP (transformer_gan.py:203-230) with a frozen encoder.  With x = x^ E (the embedded interpolate the
        reference differentiates with respect to, :211-216), h0 = Enc(x)[:, 0], D = head_theta(h0):
            g = grad_x D = J^T a,   a = d head / d h0  (a function of theta),   J = d h0 / d x (no theta inside)
            gp = 10 mean_b (||g_b|| - 1)^2
        d gp / d theta flows only through a: _VjpFn's backward hands autograd J w (one forward-tangent pass of the
        encoder) and torch differentiates a(theta) on [B, hidden] tensors.  No double-backward graph of the encoder."""
        from tgan_b200 import bert as TB
        B, T, _ = fake_bt.shape
        if self.gp_alpha_source is not None:
            alpha = self.gp_alpha_source(B).to(device=fake_bt.device, dtype=fake_bt.dtype).view(B, 1, 1)
        else:
            alpha = torch.rand([B, 1, 1], device=fake_bt.device, dtype=fake_bt.dtype)
        xhat = (alpha * real_1h + (1 - alpha) * fake_bt).detach()
        c = eng.forward(B, T, x=eng.embed_onehot(xhat), training=self.discriminator.training, save=True)
        h0 = c.h0.detach().requires_grad_(True)
        with torch.enable_grad():
            d = self._bert_head(h0)
            (a,) = torch.autograd.grad(d, h0, grad_outputs=torch.ones_like(d), create_graph=True)
            g = TB._VjpFn.apply(eng, c, a)
            slopes = torch.sqrt(g.reshape(B, -1).pow(2).sum(1) + 1e-12)
            return ((slopes - 1.0) ** 2).mean() * LAMBDA

    def calc_gradient_penalty(self, real_data, fake_data, LAMBDA=10):
        """WGAN-GP on interpolated one-hot rows (transformer_gan.py:203-230).  real / fake: [B, T, V'] float."""
        B = real_data.shape[0]
        if self.gp_alpha_source is not None:
            alpha = self.gp_alpha_source(B).to(device=real_data.device, dtype=real_data.dtype).view(B, 1, 1)
        else:
            alpha = torch.rand([B, 1, 1], device=real_data.device, dtype=real_data.dtype)
        x = (alpha * real_data + (1 - alpha) * fake_data).detach().requires_grad_(True)
        if self.cfg.DISCRIMINATOR.type == "bert":
            # the reference re-binds ``interpolates`` to the embedded rows before differentiating (:211-216): the
            # penalised gradient is the one w.r.t. x^ E_bert ([B, T, hidden]), not w.r.t. the one-hot interpolate
            x = torch.einsum("ve,bcv->bce", self._bert_embedding_matrix(), x)
            d = self._bert_logit(x)
        else:
            d = self.discriminator(x)
        (grad,) = torch.autograd.grad(d, x, grad_outputs=torch.ones_like(d), create_graph=True, retain_graph=True)
        slopes = torch.sqrt(grad.reshape(B, -1).pow(2).sum(1) + 1e-12)
        return ((slopes - 1.0) ** 2).mean() * LAMBDA

    # ------------------------------------------------------------------------------------------------ sampling
    def _one_hot(self, ids):
        return torch.zeros(*ids.shape, self.ntokens, dtype=torch.float32, device=ids.device).scatter_(-1, ids[..., None], 1.0)

    def _sample_chunks(self, data, with_grad):
        """Yields ``(chunk_start, chunk_end, fake_chunk [len, B, V])`` for each of the ``sample_chunks_mem`` pieces of
        the ``DISCRIMINATOR.tgt_len`` sequence (transformer_gan.py:273-349).  The first chunk starts with the
        ``context_len`` real tokens as one-hot rows; the gradient chain inside a chunk is the soft one-hot fed back as
        the next input, and each chunk's first generated token restarts from a hard id.
        (Measured and dropped: cutting the batch into column lanes on separate streams.  The chain is a string of
        DEPENDENT launch-latency-bound kernels; halving the rows of each does not shorten any of them, so two lanes
        side by side take as long as one -- 245 vs 238 ms per generator update at 512 sequences.)"""
        dcfg, gen = self.cfg.DISCRIMINATOR, self.generator
        mems = None
        if dcfg.context_len > 1:
            with torch.no_grad():
                _, mems = gen.forward_generate(data[:dcfg.context_len - 1], mems)
        chunk = dcfg.tgt_len // dcfg.sample_chunks_mem
        seq, ids, step = [], [], 0
        for cs in range(0, dcfg.tgt_len, chunk):
            ce = min(cs + chunk, dcfg.tgt_len)
            for pos in range(cs, ce):
                if pos < dcfg.context_len:
                    seq.append(self._one_hot(data[pos]))
                    continue
                prev = seq[-1]
                hard = prev.argmax(dim=-1)[None, :].detach()
                inp = hard if (dcfg.truncate_backprop or pos == cs or not with_grad) else prev[None]
                noise = None if self.gumbel_noise_source is None else self.gumbel_noise_source(step, (1,) + tuple(prev.shape))
                st, mems = gen.forward_generate_gumbel(inp, self.temperature, mems, noise=noise)
                seq.append(st[0])
                ids.append(st[0].detach().argmax(dim=-1))
                step += 1
            if len(seq) == chunk + 1:  # later chunks carry the previous chunk's last token only as the seed
                seq = seq[1:]
            yield cs, ce, torch.stack(seq, 0)
            seq = [seq[-1].detach()]
        self.last_sampled_ids = torch.stack(ids, 0) if ids else None

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, data, target, reset_mems, train_loss, mems=None, status_vec=None, update_D0=False):
        out = {"mle": None, "gen_loss": None, "dis_loss": None, "mems": None}
        if status_vec is not None or self.cfg.TRAIN.append_note_status:
            raise NotImplementedError("append_note_status is off in every shipped config and not accelerated")
        if "classifier" in train_loss and not self.ppo:
            raise AttributeError("'classifier' phase without a PPO loss type: the model has no dis_D")
        if "mle" in train_loss:
            out["mle"], out["mems"] = self.generator(data, target, reset_mems, mems)
        if "gen" not in train_loss and "dis" not in train_loss and "classifier" not in train_loss:
            return out
        self._update_D0 = bool(update_D0)
        # In bf16 mode the discriminator's GEMMs (HuggingFace BERT / RelGAN_D: library code, fp32 tensors) run on the
        # tensor cores as TF32 -- more mantissa than the generator's own bf16 operands; fp32 mode keeps them exact.
        if isinstance(self.discriminator, RelGAN_D):
            # bf16 mode without a gradient penalty (its double backward needs torch's differentiable backward)
            self.discriminator.own_gemm = bool(self.use_own_cnn_gemm and gen_is_bf16(self.generator) and
                                               "gp" not in self.cfg.DISCRIMINATOR.CNN.loss_type)
        tf32 = self.disc_tf32 and gen_is_bf16(self.generator)
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32 or prev_tf32
        try:
            injected = self.gumbel_noise_source is not None or self.gp_alpha_source is not None
            # the PPO variants carry host-side state between calls (P0, update_D0): launched from the host
            if (self.use_cuda_graphs and data.is_cuda and (not injected or self.sources_graph_safe)
                    and self.cfg.DISCRIMINATOR.backprop_outside and not self.ppo):
                out.update(self._gan_phase_graphed(data, train_loss))
            else:
                out.update(self._gan_phase(data, train_loss))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        return out

    # ------------------------------------------------------------------------------------------------ CUDA graphs
    def _grad_tensors(self):
        """Every tensor the GAN phase accumulates gradients into (created as zeros if missing): their addresses are
        part of a captured graph."""
        ts = []
        for m in (self.generator, self.discriminator):
            for p in m.parameters():
                if p.requires_grad:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    ts.append(p.grad)
        return ts

    def _gan_phase_graphed(self, data, train_loss):
        """The whole adversarial phase -- context pass, the 123-step Gumbel sampling chain, discriminator forward /
        backward [/ gradient penalty], generator backward through the chain -- as ONE CUDA-graph replay.  The phase is
        ~7 000 (dis) / ~23 000 (gen) launches of tiny kernels whose shapes depend only on (phase, batch): the host
        cannot enqueue them as fast as the GPU retires them.  The first call of a key runs eagerly (lazy
        initialisation), the second is captured, later ones are replayed.  The annealed temperature is read from a
        device scalar, the noise / dropout streams advance through the device step counter."""
        from tgan_b200 import lib as L
        gen = self.generator
        eng = gen._get_engine()
        key = (train_loss, tuple(data.shape), gen.training, self.discriminator.training, eng._param_key)
        if key not in self._gan_graphs:
            if key not in self._gan_warm:
                self._gan_warm.add(key)
                return self._gan_phase(data, train_loss)
            grads = self._grad_tensors()
            names = [r for r, *_ in eng.layout.reference_map()]
            pd = dict(gen.named_parameters())
            pd.setdefault("crit.out_layers.0.weight", gen.crit.out_layers[0].weight)
            entry = type("GanGraph", (), {})()
            # descriptor table staged outside the capture; the entry owns it for the graph's lifetime (the captured
            # unpack kernel reads it at every replay -- the engine's cache may evict its own reference)
            entry.desc = eng._unpack_desc_for({n: pd[n].grad for n in names})
            entry.data, entry.tau = data.clone(), torch.ones(1, dtype=torch.float32, device=data.device)
            entry.grad_ptrs = tuple(t.data_ptr() for t in grads)
            ctr = L.step_counter(data.device)
            eng.invalidate()  # the parameter re-pack must be part of the graph
            saved_tau, self.temperature = self.temperature, entry.tau
            import gc
            gc.collect()  # no autograd graph of the eager warm-up call (built on another stream) may survive into the capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            try:
                with torch.cuda.graph(g, pool=eng.graph_pool()):
                    ctr.add_(1)
                    entry.out = self._gan_phase(entry.data, train_loss)
            finally:
                self.temperature = saved_tau
                eng.invalidate()
            entry.graph, entry.n = g, L.launch_count() - n0
            self._gan_graphs[key] = entry
        entry = self._gan_graphs[key]
        if tuple(t.data_ptr() for t in self._grad_tensors()) != entry.grad_ptrs:
            del self._gan_graphs[key]  # the gradient buffers moved (zero_grad(set_to_none=True)): capture again
            return self._gan_phase_graphed(data, train_loss)
        entry.data.copy_(data)
        entry.tau.fill_(float(self.temperature))
        entry.graph.replay()
        L.note_graph_replay(entry.n)
        eng.pack_epoch += 1
        return {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in entry.out.items()}

    def _gan_phase(self, data, train_loss):
        """Adversarial part of ``forward`` (transformer_gan.py:273-533): returns the dis / gen / gp entries of the dict."""
        out = {}
        dcfg, gen = self.cfg.DISCRIMINATOR, self.generator
        dtype_cfg = dcfg.BERT if dcfg.type == "bert" else dcfg.CNN
        if dcfg.type not in ("bert", "cnn"):
            raise NotImplementedError(dcfg.type)
        train_dis = "dis" in train_loss
        train_gen = "gen" in train_loss and not train_dis
        train_cls = "classifier" in train_loss
        ppo_gen = train_gen and self.ppo  # :375 (either discriminator's loss type switches the ratio on)
        use_gp = train_dis and "gp" in dtype_cfg.loss_type
        share = dcfg.batch_chunk * dcfg.sample_chunks_mem
        cached = (gen.tgt_len, gen.mem_len)
        gen.reset_length(1, dcfg.mem_len)
        gen.detach_mems_grad = False
        g_total = d_total = gp_total = 0
        try:
            sampler = self._sample_chunks(data, with_grad=train_gen and not train_cls)  # 'classifier': hard ids, :311
            while True:
                if train_gen:
                    item = next(sampler, None)
                else:
                    with torch.no_grad():
                        item = next(sampler, None)
                if item is None:
                    break
                cs, ce, fake = item
                if train_dis:
                    fake = fake.detach()
                if train_cls:  # BCE update of the density-ratio classifier on real / sampled chunks (:350-372)
                    if self.P0 is None:
                        self.P0 = self._ppo_P0(fake)
                    n = self.P0.shape[0]
                    bce = torch.nn.functional.binary_cross_entropy
                    err = bce(torch.sigmoid(self.dis_D_forward(data[cs:ce])), self.P0.new_ones(n)) + \
                        bce(torch.sigmoid(self.dis_D_forward(fake.detach())), self.P0.new_zeros(n))
                    (err.float().mean() / share).backward()
                    continue
                ratio = None
                if ppo_gen:  # clipped density ratio of the current samples under dis_D (:375-388)
                    if self.P0 is None or self._update_D0:
                        self.P0 = self._ppo_P0(fake)
                    D1 = torch.sigmoid(self.dis_D_forward(fake))
                    ratio = (1.0 - D1) / torch.clamp(D1 * self.P0, min=1e-7)
                    ratio_clipped = torch.clamp(ratio, 1.0 - self.cfg.PPO.clip_param, 1.0 + self.cfg.PPO.clip_param)
                real = data[cs:ce].transpose(0, 1)          # [B, len] ids
                fake_bt = fake.transpose(0, 1)              # [B, len, V]
                if dcfg.type == "bert":
                    # BERT's vocabulary has one extra (MASK) id: pad the sampled rows with a zero column (:396-399)
                    fake_bt = torch.cat([fake_bt, fake_bt.new_zeros(*fake_bt.shape[:-1], 1)], -1)
                    own = self._own_bert()
                    if own is not None:
                        d_real = self._own_logit(own, ids=real)
                        d_fake = self._own_logit(own, soft=fake_bt)
                    else:
                        E = self._bert_embedding_matrix()
                        d_real = self._bert_logit(E[real])
                        d_fake = self._bert_logit(torch.einsum("ve,bcv->bce", E, fake_bt))
                    real_1h = None
                    if use_gp:
                        real_1h = torch.zeros(*real.shape, self.ntokens + 1, device=real.device).scatter_(-1, real[..., None], 1.0)
                else:
                    real_1h = self._one_hot(real)
                    d_real = self.discriminator(real_1h)
                    d_fake = self.discriminator(fake_bt)
                if ratio is not None and "ppo" in dtype_cfg.loss_type:  # PPO surrogate target (:419-424, :456-461)
                    surr1, surr2 = ratio * d_fake, ratio_clipped * d_fake
                    d_fake = torch.where(d_fake > 0, torch.min(surr1, surr2), torch.max(surr1, surr2))
                g_loss, d_loss = get_losses(d_real, d_fake, dtype_cfg.loss_type)
                if use_gp and dcfg.type == "bert" and self._own_bert() is not None:
                    gp = self._own_gradient_penalty(self._own_bert(), real_1h, fake_bt)
                else:
                    gp = self.calc_gradient_penalty(real_1h, fake_bt) if use_gp else None
                keep = (lambda t: t.detach()) if dcfg.backprop_outside else (lambda t: t)
                g_total = g_total + keep(g_loss)
                d_total = d_total + keep(d_loss)
                if gp is not None:
                    gp_total = gp_total + keep(gp)
                if dcfg.backprop_outside:  # the reference's name for "backward happens here, inside forward" (:487-502)
                    if train_gen or ("gen" in train_loss and not train_dis):
                        with gen.grad_window():  # the chunk's single-token backward calls share one zero-fill / unpack
                            (g_loss.float().mean() * dcfg.gen_loss_factor / share).backward()
                    if train_dis:
                        (d_loss.float().mean() * dcfg.dis_loss_factor / share).backward()
                        if gp is not None:
                            (gp.float().mean() * dcfg.dis_loss_factor / share).backward()
        finally:
            gen.detach_mems_grad = True
            gen.reset_length(*cached)
        if train_cls:
            return out
        if train_dis:
            out["dis_loss"] = dcfg.dis_loss_factor * d_total / dcfg.sample_chunks_mem
            if use_gp:
                out["gp_loss"] = dcfg.dis_loss_factor * gp_total / dcfg.sample_chunks_mem
        else:
            out["gen_loss"] = dcfg.gen_loss_factor * g_total / dcfg.sample_chunks_mem
        return out
