"""Kernel timeline of ONE graph replay of an adversarial update (CUPTI through torch.profiler -- kernels inside a graph
launch are reported individually): per-kernel totals, busy time (union of the kernel intervals over all streams),
idle gaps, and how much runs concurrently.   python tools/graph_timeline.py [gen_loss|dis_loss] [B] [dump.csv]"""
import collections
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

phase = sys.argv[1] if len(sys.argv) > 1 else "gen_loss"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dump = sys.argv[3] if len(sys.argv) > 3 else None
args = types.SimpleNamespace(workload="gan", global_batch=B, scaling="weak", batch_chunk=1, dtype="bf16", kernel_impl=0,
                             no_graphs=False, no_buckets=True)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cyc = bench.Cycle(args, dev, 1, 0)
cyc.mle_step(False)
opt = cyc.dis_opt if phase == "dis_loss" else cyc.gen_opt


def call():
    cyc.model(cyc.dev_dis[0], None, None, phase)
    opt.step()
    cyc.fp.zero_grad()
    cyc.dfp.zero_grad()


for _ in range(4):
    call()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    call()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0 and
      "memcpy" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
rows = [(e.time_range.start - t0, e.time_range.end - t0, e.name) for e in ev]
span = max(r[1] for r in rows)
busy, cur_s, cur_e, conc = 0.0, None, None, 0.0
for s, e, _ in rows:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        conc += min(e, cur_e) - s
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
tot = sum(e - s for s, e, _ in rows)
print(f"{phase} B={B}: {len(rows)} kernels, span {span / 1e3:.2f} ms, busy (union) {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms, "
      f"sum of kernel durations {tot / 1e3:.2f} ms (overlapped {100 * (tot - busy) / tot:.1f} %)")
def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "")
    return n.split("(")[0][:90]


agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in rows:
    n = short(n)
    agg[n][0] += 1
    agg[n][1] += e - s
for n, v in sorted(agg.items(), key=lambda x: -x[1][1])[:32]:
    print(f"{v[1] / 1e3:9.2f} ms {100 * v[1] / tot:5.1f}% {v[0]:6d} x {v[1] / v[0]:7.1f} us  {n}")
if dump:
    with open(dump, "w") as f:
        f.write("start_us,end_us,name\n")
        for s, e, n in rows:
            f.write(f"{s:.2f},{e:.2f},{short(n)}\n")
