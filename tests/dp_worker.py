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
    first_grad = None
    for s in range(data.shape[0]):
        d, t = data[s][:, cols].contiguous().to(dev), target[s][:, cols].contiguous().to(dev)
        loss, mems = model(d, t, None, mems)
        (loss.mean() * scale).backward()
        if s == 0:  # gradient of the first step (identical parameters in every mode), summed over the ranks
            first_grad = fp.grad.clone()
            if mode == "flat":
                dist.all_reduce(first_grad)
            elif mode == "single":
                first_grad *= world  # mean over the whole batch = average of the shard means: x world = their sum
        opt.step()
    torch.cuda.synchronize()
    return fp.flat.clone(), (0 if reducer is None else reducer.buckets), first_grad


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
        flat, _, g_flat = run("flat", graphs, world, rank, dev, data, target)
        flat2, _, g_flat2 = run("flat", graphs, world, rank, dev, data, target)
        buck, nb, g_buck = run("bucket", graphs, world, rank, dev, data, target)
        single, _, g_single = run("single", graphs, world, rank, dev, data, target)
        rel = lambda a, b: ((a - b).norm() / b.norm()).item()
        # the exchanged gradient itself: bucketed == flat to fp32 rounding; both == the single-process gradient up to the
        # bf16 effects of a different batch split (GEMM row counts / split-K partitions differ)
        gb, gs, noise = rel(g_buck, g_flat), rel(g_flat, g_single), rel(g_flat2, g_flat)
        # parameters after 6 clipped Adam steps: Adam divides by sqrt(v), so rounding-level gradient differences of
        # near-zero entries grow; the run-to-run floor (flat vs flat: fp32 atomics order) is the yardstick
        p_noise, p_buck = (flat2 - flat).abs().max().item(), (buck - flat).abs().max().item()
        p_single = rel(flat, single)
        stats = torch.tensor([gb, gs, noise, p_noise, p_buck, p_single], device=dev)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        gb, gs, noise, p_noise, p_buck, p_single = stats.tolist()
        if rank == 0:
            print(f"graphs={graphs} buckets={nb} grad: |bucket-flat|={gb:.2e} |flat-flat'|={noise:.2e} |flat-single|={gs:.2e}  "
                  f"params: |flat-flat'|={p_noise:.2e} |bucket-flat|={p_buck:.2e} rel |flat-single|={p_single:.2e}", flush=True)
        ok = ok and nb > 0 and gb <= max(5 * noise, 1e-5) and gs < 2e-2 and p_buck <= max(5 * p_noise, 1e-5) and p_single < 0.1
    dist.destroy_process_group()
    if rank == 0:
        print("DP_WORKER_OK" if ok else "DP_WORKER_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
