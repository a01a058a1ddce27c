"""CPU, world_size 2 (gloo): the host-side data-parallel logic -- batch-column sharding, flat parameter / gradient
buffers and the once-per-step gradient all-reduce -- reproduces the single-process gradient of the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _model(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 5))


def _batch():
    g = torch.Generator().manual_seed(123)
    return torch.randn(9, 8, 12, generator=g), torch.randn(9, 8, 5, generator=g)  # [len, batch columns, feat]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tgan_b200 import dp
    model = _model()
    fp = dp.FlatParams(model.parameters())
    x, y = _batch()
    lo, hi = dp.shard_columns(x.shape[1], world, rank)
    # two micro-batches per rank, accumulated locally; the all-reduce runs ONCE (train.py's DDP would run it twice)
    mid = (lo + hi) // 2
    for a, b in ((lo, mid), (mid, hi)):
        loss = ((model(x[:, a:b]) - y[:, a:b]) ** 2).sum() / (x.shape[0] * x.shape[1])
        loss.backward()
    dp.allreduce_gradients(fp.grad, world)
    q.put((rank, fp.grad.clone(), dp.rank_seed(1111, rank)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gradient_equals_single_process():
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "transformer-gan_b200")
    os.environ["PYTHONPATH"] = pkg + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = _model()
    x, y = _batch()
    (((model(x) - y) ** 2).sum() / (x.shape[0] * x.shape[1])).backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    for rank, g, seed in res:
        assert torch.allclose(g, want, rtol=1e-5, atol=1e-6), (rank, (g - want).abs().max())
        assert seed == 1111 + 1000 * rank


def test_flat_params_are_views_and_shards_partition_the_batch():
    from tgan_b200 import dp
    model = _model(1)
    before = [p.detach().clone() for p in model.parameters()]
    fp = dp.FlatParams(model.parameters())
    assert fp.numel() == sum(p.numel() for p in model.parameters())
    for p, b, (off, k) in zip(model.parameters(), before, fp.slices):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == fp.flat[off:off + k].data_ptr() and p.grad.data_ptr() == fp.grad[off:off + k].data_ptr()
    cols = [dp.shard_columns(512, 8, r) for r in range(8)]
    assert cols[0] == (0, 64) and cols[-1] == (448, 512) and all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
    with pytest.raises(ValueError):
        dp.shard_columns(10, 4, 0)
