"""Diagnosis only: torch.profiler kernel-time table of decode steps (B=128, memory 4146)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
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
    for _ in range(3):
        _, mems = model.forward_generate(tok, mems)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(8):
            _, mems = model.forward_generate(tok, mems)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
