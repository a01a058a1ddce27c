"""CNN discriminators of the GAN step, with the reference's class names and state_dict keys.

Reference: ``CNNDiscriminator`` (model/discriminator.py:27-82) and ``RelGAN_D`` (model/transformer_gan.py:44-119), the
discriminator of ``experiment_cnn.yml``.  Its input is the generator's straight-through one-hot sequence
``[batch, seq, vocab]`` written by the fused Gumbel sampler; the body (a bias-free embedding projection, four
time-convolutions over ``num_rep`` independent representations, max-over-time, highway, two linears): in bf16 mode
its three dense layers (98 % of the FLOPs: embedding projection, the 1200 x 1200 highway layer, feature2out) run on the
library's tcgen05 GEMM through ``tgan_b200.nn.linear``; the 2..5-tap time convolutions (contraction length <= 5: not
tensor-core work), ReLU / max-over-time and the 100 -> 1 output layer are torch.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from utils.helpers import truncated_normal_

dis_filter_sizes = [2, 3, 4, 5]
dis_num_filters = [300, 300, 300, 300]


class CNNDiscriminator(nn.Module):
    def __init__(self, embed_dim, vocab_size, filter_sizes, num_filters, padding_idx, gpu=False, dropout=0.2, cfg=None):
        super().__init__()
        self.embedding_dim, self.vocab_size, self.padding_idx = embed_dim, vocab_size, padding_idx
        self.feature_dim = sum(num_filters)
        self.gpu, self.cfg = gpu, cfg
        self.embeddings = nn.Embedding(vocab_size, embed_dim, padding_idx=padding_idx)
        self.convs = nn.ModuleList([nn.Conv2d(1, n, (f, embed_dim)) for n, f in zip(num_filters, filter_sizes)])
        self.highway = nn.Linear(self.feature_dim, self.feature_dim)
        self.feature2out = nn.Linear(self.feature_dim, 2)
        self.dropout = nn.Dropout(dropout)
        self.init_params()

    @staticmethod
    def _highway(gate_in, x):
        g = torch.sigmoid(gate_in)
        return g * F.relu(gate_in) + (1.0 - g) * x

    def get_feature(self, inp):
        emb = self.embeddings(inp).unsqueeze(1)  # [B, 1, T, E]
        pooled = [F.relu(conv(emb)).squeeze(3).amax(dim=2) for conv in self.convs]
        feat = torch.cat(pooled, 1)
        return self._highway(self.highway(feat), feat)

    def forward(self, inp):
        return self.feature2out(self.dropout(self.get_feature(inp)))

    def init_params(self):
        """discriminator.py:73-82: every tensor with >= 1 dim, std = 1/sqrt(shape[0])."""
        mode = self.cfg.DISCRIMINATOR.CNN.init if self.cfg is not None else "uniform"
        for param in self.parameters():
            if param.requires_grad and param.dim() > 0:
                std = 1 / math.sqrt(param.shape[0])
                if mode == "uniform":
                    nn.init.uniform_(param, a=-0.05, b=0.05)
                elif mode == "normal":
                    nn.init.normal_(param, std=std)
                elif mode == "truncated_normal":
                    truncated_normal_(param, std=std)


class RelGAN_D(CNNDiscriminator):
    """transformer_gan.py:44-119.  forward(inp [B, T, vocab] float) -> logits [B * num_rep]."""

    def __init__(self, embed_dim, max_seq_len, num_rep, vocab_size, padding_idx, gpu=True, dropout=0.25, cfg=None):
        super().__init__(embed_dim, vocab_size, dis_filter_sizes, dis_num_filters, padding_idx, gpu, dropout, cfg)
        self.embed_dim, self.max_seq_len = embed_dim, max_seq_len
        self.feature_dim = sum(dis_num_filters)
        self.emb_dim_single = int(embed_dim / num_rep)
        self.embeddings = nn.Linear(vocab_size, embed_dim, bias=False)
        self.convs = nn.ModuleList([
            nn.Conv2d(1, n, (f, self.emb_dim_single), stride=(1, self.emb_dim_single))
            for n, f in zip(dis_num_filters, dis_filter_sizes)])
        self.highway = nn.Linear(self.feature_dim, self.feature_dim)
        self.feature2out = nn.Linear(self.feature_dim, 100)
        self.out2logits = nn.Linear(100, 1)
        self.dropout = nn.Dropout(dropout)
        self.init_params()

    # set by TransformerGAN in bf16 mode when no gradient penalty (= no double backward) is configured: the three
    # dense layers run on the library's tcgen05 GEMM (tgan_b200.nn.linear) instead of cuBLAS
    own_gemm = False

    def _lin(self, layer, x):
        if self.own_gemm and x.is_cuda:
            from tgan_b200.nn import linear
            return linear(x, layer.weight, layer.bias)
        return layer(x)

    def forward(self, inp):
        B, T, V = inp.shape
        emb = self._lin(self.embeddings, inp.reshape(B * T, V)).view(B, 1, T, self.embed_dim)  # [B, 1, T, embed_dim]
        pooled = [F.relu(conv(emb)).amax(dim=2) for conv in self.convs]  # each [B, filters, num_rep]
        feat = torch.cat(pooled, 1).permute(0, 2, 1).reshape(-1, self.feature_dim)  # [(B * num_rep), feature_dim]
        feat = self._highway(self._lin(self.highway, feat), feat)
        return self.out2logits(self._lin(self.feature2out, self.dropout(feat))).squeeze(1)
