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


class Lamb(torch.optim.Optimizer):
    """Drop-in for the reference's ``lamb.Lamb`` (lamb.py:21-118: same constructor, ``param_groups``, ``state`` entries
    ``step`` / ``exp_avg`` / ``exp_avg_sq`` / ``weight_norm`` / ``adam_norm`` / ``trust_ratio``) whose ``step()`` is the
    fused kernel pair (``tgan_lamb_step``) instead of ~10 small torch kernels per tensor.

    A trainer keeps its own objects -- ``optimizer = Lamb(model.generator.parameters(), lr=..)``, ``clip_grad_norm_``,
    LR schedulers writing ``param_groups[i]['lr']``, ``zero_grad()``, ``state_dict()`` -- (train.py:396-398, 914-921).  On
    the first step the parameters of a group that carry a gradient are moved into one flat buffer (they become views of
    it, as with ``FlatParams``); the moments live in flat buffers too and ``state[p]['exp_avg']`` etc. are views of them.
    Gradients are gathered with one multi-tensor copy per step (``zero_grad()`` drops the ``.grad`` tensors, so they
    cannot be aliased permanently).  Parameters without a gradient are skipped exactly as the reference skips them.
    To run the unmodified train.py on it: ``sys.modules['lamb'] = tgan_b200.dp`` before train.py imports ``lamb``.
    CUDA only."""

    CHUNK = FusedLamb.CHUNK

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0, adam=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {}".format(betas))
        self.adam = adam
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = {}  # group index -> dict(key, params, flat, grad, m, v, upd, chunks, norms, slices)

    def _build(self, gi, ps):
        from . import lib as L
        dev = ps[0].device
        if dev.type != "cuda":
            raise L.TganError("tgan_b200.dp.Lamb runs on CUDA only (no host fallback)")
        n = sum(p.numel() for p in ps)
        f = dict(key=tuple(id(p) for p in ps), params=ps, flat=torch.empty(n, dtype=torch.float32, device=dev),
                 grad=torch.empty(n, dtype=torch.float32, device=dev), m=torch.zeros(n, dtype=torch.float32, device=dev),
                 v=torch.zeros(n, dtype=torch.float32, device=dev), upd=torch.empty(n, dtype=torch.float32, device=dev))
        off, rows, f["slices"] = 0, [], []
        for tid, p in enumerate(ps):
            k = p.numel()
            f["flat"][off:off + k].copy_(p.data.reshape(-1))
            p.data = f["flat"][off:off + k].view_as(p)
            st = self.state[p]
            for name, buf in (("exp_avg", f["m"]), ("exp_avg_sq", f["v"])):
                view = buf[off:off + k].view_as(p)
                if name in st:  # moments from an earlier layout / a loaded state_dict are carried over
                    view.copy_(st[name])
                st[name] = view
            st.setdefault("step", 0)
            f["slices"].append((off, k))
            for c0 in range(0, k, self.CHUNK):
                rows.append([tid, off + c0, min(self.CHUNK, k - c0)])
            off += k
        f["chunks"] = torch.tensor(rows, dtype=torch.int64, device=dev)
        f["norms"] = torch.zeros(2 * len(ps), dtype=torch.float32, device=dev)
        f["grad_views"] = [f["grad"][o:o + k].view_as(p) for p, (o, k) in zip(ps, f["slices"])]
        self._flat[gi] = f
        return f

    @torch.no_grad()
    def step(self, closure=None):
        from . import lib as L
        from .engine import notify_params_updated
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            f = self._flat.get(gi)
            if f is None or f["key"] != tuple(id(p) for p in ps) or \
                    any(self.state[p].get("exp_avg") is None or
                        self.state[p]["exp_avg"].data_ptr() != f["m"].data_ptr() + 4 * o
                        for p, (o, _) in zip(ps, f["slices"])):
                f = self._build(gi, ps)  # first step, a changed set of trained tensors, or a loaded state_dict
            torch._foreach_copy_(f["grad_views"], [p.grad for p in ps])
            L.lamb_step(f["flat"], f["grad"], f["m"], f["v"], f["upd"], f["chunks"], f["chunks"].shape[0], f["norms"],
                        len(ps), group["lr"], group["betas"][0], group["betas"][1], group["eps"], group["weight_decay"],
                        None, 0.0, 1.0, adam=self.adam)
            nr = f["norms"].view(-1, 2).sqrt()
            wn, an = nr[:, 0].clamp(0, 10), nr[:, 1]
            tr = torch.where((wn == 0) | (an == 0), torch.ones_like(wn), wn / (an + group["eps"]))
            for i, p in enumerate(ps):  # device scalars (views): no host synchronisation
                st = self.state[p]
                st["step"] += 1
                st["weight_norm"], st["adam_norm"], st["trust_ratio"] = wn[i], an[i], tr[i]
        notify_params_updated()
        return loss
