"""torch.profiler kernel table of one adversarial phase of bench.py's cycle (in situ: real cache state, host launches).
Usage: python tools/cycle_prof.py [dis_loss|gen_loss|mle] [B]"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
phase = sys.argv[1] if len(sys.argv) > 1 else "gen_loss"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sys.argv = ["bench.py", "--no-graphs", "--global-batch", str(B)]
import bench
import torch
from torch.profiler import profile, ProfilerActivity
args = bench.parse()
dev = torch.device("cuda", 0)
cyc = bench.Cycle(args, dev, 1, 0)
if phase == "mle":
    for _ in range(10):
        cyc.mle_step(False)
    fn = lambda: cyc.mle_step(False)
else:
    opt = cyc.dis_opt if phase == "dis_loss" else cyc.gen_opt
    def fn():
        cyc.model(cyc.dev_dis[0], None, None, phase)
        opt.step(); cyc.fp.zero_grad(); cyc.dfp.zero_grad()
    fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn()
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total, e.count, e.key) for e in ev if e.device_time_total > 0), reverse=True)
tot = sum(r[0] for r in rows)
print(f"{phase} B={B}: total kernel time {tot/1e3:.1f} ms over {sum(r[1] for r in rows)} kernels")
for t, n, k in rows[:40]:
    print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}% {n:6d} x {t/n:8.1f} us  {k[:110]}")
