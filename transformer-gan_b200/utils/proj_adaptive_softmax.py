"""Parameter container with the reference's class name and state_dict keys for the output layer.

Reference: model/utils/proj_adaptive_softmax.py:6-148.  With ``cutoffs == []`` (always, mem_transformer.py:407-409)
the reference layer is ``Linear(d_embed -> n_token)`` + ``log_softmax`` + ``gather``; here the arithmetic is done by
tgan_gemm (logits, bias epilogue) + tgan_ce_fwd / tgan_ce_bwd inside the generator's engine, so this module only
owns ``out_layers.0.{weight,bias}`` / ``out_projs`` for ``weights_init`` (train.py:313-363) and checkpoints.
The adaptive-cluster branch (:85-146) is dead code for every shipped config and is not provided.
"""
import torch
import torch.nn as nn


class ProjectedAdaptiveLogSoftmax(nn.Module):
    def __init__(self, n_token, d_embed, d_proj, cutoffs=[], div_val=1, keep_order=False):
        super().__init__()
        if cutoffs or div_val != 1:
            raise NotImplementedError("adaptive-softmax clusters are not part of the accelerated path")
        if d_proj != d_embed:
            raise NotImplementedError("d_proj != d_embed is not part of the accelerated path")
        self.n_token, self.d_embed, self.d_proj = n_token, d_embed, d_proj
        self.cutoffs = [n_token]
        self.cutoff_ends = [0] + self.cutoffs
        self.div_val = div_val
        self.shortlist_size = n_token
        self.n_clusters = 0
        self.head_size = n_token
        self.out_layers = nn.ModuleList([nn.Linear(d_embed, n_token)])
        self.out_projs = nn.ParameterList()
        self.out_projs.append(None)
        self.keep_order = keep_order

    def forward(self, hidden, target, keep_order=False):
        raise RuntimeError("ProjectedAdaptiveLogSoftmax is evaluated inside MemTransformerLM's fused engine "
                           "(tgan_gemm + tgan_ce_fwd); call MemTransformerLM.forward instead")
