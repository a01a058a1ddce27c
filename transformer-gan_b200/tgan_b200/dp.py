"""Data-parallel plumbing of the training step: flat parameter / gradient buffers, one gradient all-reduce per
optimizer step, fused clip + Adam.

Reference behaviour replaced (file:line relative to /root/reference/model):
  * ``DistributedDataParallel(model, ...)``            train.py:649-655 -- the reference all-reduces the generator
    gradients once per MICRO-batch (no ``no_sync``) and never syncs the in-forward GAN gradients;
    here every rank accumulates its micro-batches locally and ONE ``all_reduce`` over the flat buffer runs per
    optimizer step (NCCL over NVLink on GPUs, gloo in the CPU tests).  The result equals the single-process run on
    the concatenated batch.
  * ``clip_grad_norm_`` + ``optimizer.step()``         train.py:914-921 -- ``tgan_sumsq`` + ``tgan_adam_step`` over the
    flat buffers (the 1/world average is folded into the Adam kernel's gradient scale).
  * per-rank data seed ``seed + 1000 * rank``          train.py:224;  lr / num_gpus  train.py:392.
The collective is a separate step, not fused into a compute kernel: no kernel of this path produces data that a
peer consumes tile by tile (the exchange is the whole 13.7 M-element gradient, once per step).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch


def unique_params(params: Iterable[torch.nn.Parameter]) -> List[torch.nn.Parameter]:
    seen, out = set(), []
    for p in params:
        if id(p) not in seen:
            seen.add(id(p))
            out.append(p)
    return out


class FlatParams:
    """Moves every parameter of ``module`` into ONE flat fp32 buffer (the parameters become views of it) with a
    matching flat gradient buffer (``p.grad`` are views), so that the all-reduce, the norm and the optimizer step are
    one call each regardless of how many tensors the model has (72 for the 6-layer generator)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = unique_params(params)
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        off = 0
        self.slices: List[Tuple[int, int]] = []
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            self.slices.append((off, k))
            off += k

    def numel(self) -> int:
        return self.flat.numel()

    def zero_grad(self):
        self.grad.zero_()


def shard_columns(n_cols: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) batch columns owned by ``rank`` (global batch fixed; train.py:226-227 uses batch_size // num_gpus)."""
    if n_cols % world:
        raise ValueError(f"global batch {n_cols} must divide by the world size {world}")
    per = n_cols // world
    return rank * per, (rank + 1) * per


def rank_seed(seed: int, rank: int) -> int:
    return seed + 1000 * rank  # train.py:224


def allreduce_gradients(flat_grad: torch.Tensor, world: int, average: bool = False) -> None:
    """One SUM all-reduce of the flat gradient buffer per optimizer step.  With ``average`` the 1/world factor is
    applied here; the GPU path leaves it to the Adam kernel's ``grad_scale`` instead."""
    if world <= 1:
        return
    import torch.distributed as dist
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    if average:
        flat_grad.mul_(1.0 / world)


class BucketReducer:
    """The data-parallel gradient exchange of the MLE step through the C-ABI (``tgan_allreduce_bucket``): the padded
    gradient buffers of the engine are all-reduced layer by layer, in the order the backward finishes them (last layer
    first), on a SIDE stream -- each bucket's NCCL kernel overlaps the backward of the layers below it.  NCCL
    operations are stream-ordered and capturable, so under ``use_cuda_graphs`` the buckets are nodes of the captured
    backward graph.  Replaces DistributedDataParallel (train.py:649-655), whose hooks do the same per autograd bucket;
    unlike DDP there is one exchange per optimizer step when ``batch_chunk`` is 1 (the reference re-reduces every
    micro-batch).  Pair it with ``FusedClipAdam(..., reduce=False)``."""

    def __init__(self, world: int, rank: int, device):
        from . import lib as L
        import torch.distributed as dist
        self.L, self.world = L, world
        uid = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(L.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)  # the torch process group only carries the 128-byte id
        self.comm = L.nccl_init(bytes(uid.cpu().tolist()), world, rank)
        self.stream = torch.cuda.Stream(device=device)
        self.pending = False
        self.buckets = 0

    def reduce(self, t: torch.Tensor, offset: int, count: int):
        """enqueue the in-place SUM all-reduce of t.view(-1)[offset : offset + count]; everything enqueued so far on the
        current stream is ordered before it"""
        if count <= 0:
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        self.L.allreduce_bucket(self.comm, t, count, offset, stream=self.stream.cuda_stream)
        self.pending = True
        self.buckets += 1

    def join(self):
        """the current stream waits for every bucket enqueued so far"""
        if self.pending:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.pending = False


class FusedClipAdam:
    """``clip_grad_norm_(max_norm)`` + Adam (no weight decay by default) on flat buffers through the CUDA library
    (tgan_sumsq + tgan_adam_step).  CUDA only -- there is no host fallback."""

    def __init__(self, fp: FlatParams, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 clip: float = 0.0, world: int = 1, reduce: bool = True):
        from . import lib as L
        self.L = L
        self.fp, self.lr, self.betas, self.eps, self.wd, self.clip, self.world = fp, lr, betas, eps, weight_decay, clip, world
        self.reduce = reduce  # False: the gradients arrive already summed over the ranks (BucketReducer)
        self.m, self.v = torch.zeros_like(fp.flat), torch.zeros_like(fp.flat)
        self.gnorm_sq = torch.zeros(1, device=fp.flat.device)
        self.steps = 0

    def step(self, lr: Optional[float] = None):
        fp, L = self.fp, self.L
        self.steps += 1
        if self.reduce:
            allreduce_gradients(fp.grad, self.world)
        self.gnorm_sq.zero_()
        L.sumsq(fp.grad, fp.grad.numel(), self.gnorm_sq)
        L.adam_step(fp.flat, fp.grad, self.m, self.v, fp.numel(), self.lr if lr is None else lr, self.betas[0],
                    self.betas[1], self.eps, self.wd, self.steps, self.gnorm_sq, self.clip, 1.0 / self.world)
        fp.zero_grad()
        # the kernel wrote the flat parameter buffer behind autograd's back (no version counter moved): tell every
        # engine to re-pack its kernel-private parameter copies at the next forward
        from .engine import notify_params_updated
        notify_params_updated()


class FusedLamb:
    """The reference's LAMB (``lamb.Lamb``, lamb.py:57-118: moments without bias correction, per-tensor trust ratio
    ``clamp(|w|, 0, 10) / (|adam_step| + eps)``) preceded by ``clip_grad_norm_`` (train.py:914-921), on the flat buffers of
    a ``FlatParams`` group through the CUDA library: tgan_sumsq + tgan_lamb_step (two passes over the group: moments /
    per-tensor norms, then the update).  Same call surface as ``FusedClipAdam``.  CUDA only."""

    CHUNK = 16384

    def __init__(self, fp: FlatParams, lr: float, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.0,
                 clip: float = 0.0, world: int = 1, reduce: bool = True, adam: bool = False):
        from . import lib as L
        self.L = L
        self.fp, self.lr, self.betas, self.eps, self.wd, self.clip, self.world = fp, lr, betas, eps, weight_decay, clip, world
        self.reduce, self.adam = reduce, adam
        dev = fp.flat.device
        self.m, self.v, self.upd = torch.zeros_like(fp.flat), torch.zeros_like(fp.flat), torch.empty_like(fp.flat)
        rows = []
        for tid, (off, k) in enumerate(fp.slices):
            for c0 in range(0, k, self.CHUNK):
                rows.append([tid, off + c0, min(self.CHUNK, k - c0)])
        self.chunks = torch.tensor(rows, dtype=torch.int64, device=dev)
        self.norms = torch.zeros(2 * len(fp.slices), dtype=torch.float32, device=dev)
        self.gnorm_sq = torch.zeros(1, device=dev)
        self.steps = 0

    def step(self, lr: Optional[float] = None):
        fp, L = self.fp, self.L
        self.steps += 1
        if self.reduce:
            allreduce_gradients(fp.grad, self.world)
        self.gnorm_sq.zero_()
        L.sumsq(fp.grad, fp.grad.numel(), self.gnorm_sq)
        L.lamb_step(fp.flat, fp.grad, self.m, self.v, self.upd, self.chunks, self.chunks.shape[0], self.norms,
                    len(fp.slices), self.lr if lr is None else lr, self.betas[0], self.betas[1], self.eps, self.wd,
                    self.gnorm_sq, self.clip, 1.0 / self.world, adam=self.adam)
        fp.zero_grad()
        from .engine import notify_params_updated
        notify_params_updated()

    def trust_ratios(self) -> torch.Tensor:
        """per-tensor trust ratio of the last step (the reference keeps it in ``state['trust_ratio']``)."""
        n = self.norms.view(-1, 2).sqrt()
        wn, un = n[:, 0].clamp(0, 10), n[:, 1]
        return torch.where((wn == 0) | (un == 0), torch.ones_like(wn), wn / (un + self.eps))
