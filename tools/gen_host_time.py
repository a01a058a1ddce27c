"""Diagnosis: host time vs GPU time of decode steps (B=128, memory 4146), plus a cProfile of the host side."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
import bench as BN
import mem_transformer as MT
dev = torch.device("cuda", 0)
model = MT.MemTransformerLM(BN.make_cfg(), 310, 0)
BN.init_like_train_py(model, 1111)
model = model.to(dev).eval()
B, mem_len = 128, 4146
g = torch.Generator().manual_seed(3)
with torch.no_grad():
    mems = None
    model.reset_length(128, mem_len)
    for _ in range(33):
        _, mems = model.forward_generate(torch.randint(2, 310, (128, B), generator=g).to(dev), mems)
    model.reset_length(1, mem_len)
    tok = torch.randint(2, 310, (1, B), generator=g).to(dev)
    for _ in range(5):
        _, mems = model.forward_generate(tok, mems)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(32):
        _, mems = model.forward_generate(tok, mems)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host {1e3*(t1-t0)/32:.2f} ms/step, total {1e3*(t2-t0)/32:.2f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(16):
        _, mems = model.forward_generate(tok, mems)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
