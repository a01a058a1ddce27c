"""Worker of tests/test_dp_gpu.py (one process per GPU under torch.distributed.run): the bucketed, overlapped gradient
exchange (tgan_allreduce_bucket on a side stream inside backward) must give the same parameters as one flat all-reduce
after backward, and both must equal the single-process run on the concatenated batch."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "transformer-gan_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def cfg(drop=0.0):
    ns = types.SimpleNamespace
    return ns(MODEL=ns(num_layers=2, num_heads=4, units=64, inner_size=128, dropout=drop, attention_dropout=drop,
                       tie_embedding=True, tie_proj=False, pre_lnorm=False, same_length=False, clamp_len=-1),
              TRAIN=ns(tgt_length=32, mem_length=64, pad_type="model", replace_start_with_pad=False, append_note_status=False))


def run(mode, graphs, world, rank, dev, data, target):
    import mem_transformer as MT
    from tgan_b200 import dp
    torch.manual_seed(0)
    model = MT.MemTransformerLM(cfg(), 310, 0)
    g = torch.Generator().manual_seed(1)
    for p in model.parameters():
        p.data.copy_(0.05 * torch.randn(p.shape, generator=g))
    model = model.to(dev).train()
    model.use_cuda_graphs = graphs
    fp = dp.FlatParams(model.parameters())
    reducer = None
    if mode == "bucket":
        reducer = model.grad_reducer = dp.BucketReducer(world, rank, dev)
    opt = dp.FusedClipAdam(fp, 0.01, clip=1.0, world=world if mode != "single" else 1, reduce=mode == "flat")
    B = data.shape[2]
    if mode == "single":
        cols = slice(0, B)
        scale = 1.0
    else:
        per = B // world
        cols = slice(rank * per, (rank + 1) * per)
        scale = 1.0  # each rank's loss.mean() over its shard; Adam's 1/world averages the summed gradient
    mems = None
    for s in range(data.shape[0]):
        d, t = data[s][:, cols].contiguous().to(dev), target[s][:, cols].contiguous().to(dev)
        loss, mems = model(d, t, None, mems)
        (loss.mean() * scale).backward()
        opt.step()
    torch.cuda.synchronize()
    return fp.flat.clone(), (0 if reducer is None else reducer.buckets)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(7)
    steps, Q, B = 6, 32, 4 * world
    data = torch.randint(2, 310, (steps, Q, B), generator=g)
    target = torch.randint(2, 310, (steps, Q, B), generator=g)
    ok = True
    for graphs in (False, True):
        flat, _ = run("flat", graphs, world, rank, dev, data, target)
        buck, nb = run("bucket", graphs, world, rank, dev, data, target)
        single, _ = run("single", graphs, world, rank, dev, data, target)
        e1 = (flat - buck).abs().max().item()
        e2 = (flat - single).abs().max().item() / single.abs().max().item()
        same = torch.tensor([e1], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"graphs={graphs} buckets={nb} |flat - bucket|={same.item():.3e} rel |flat - single|={e2:.3e}", flush=True)
        ok = ok and same.item() < 1e-6 and e2 < 2e-2 and nb > 0
    dist.destroy_process_group()
    if rank == 0:
        print("DP_WORKER_OK" if ok else "DP_WORKER_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
