"""Diagnosis only: torch.profiler (CUPTI) kernel-time table of one eager GAN phase.  Usage: gan_prof.py [B] [phase]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from torch.profiler import profile, ProfilerActivity
import gan_bench as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
phase = sys.argv[2] if len(sys.argv) > 2 else "gen_loss"
dev = torch.device("cuda", 0)
model = G.build(dev)
model.temperature = 1.0
data = torch.randint(2, 310, (128, B), generator=torch.Generator().manual_seed(7)).to(dev)
model(data, None, None, phase)
model.zero_grad(set_to_none=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model(data, None, None, phase)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
