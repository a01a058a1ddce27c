"""Per-kernel shares of the benchmarked cycle from two ncu launch lists made with tools/prof_cycle.py:
    python tools/cycle_summary.py launches_mle.csv launches_gan.csv
cycle = 5 x (MLE step) + 1 dis update + 1 gen update (DISCRIMINATOR.dis_loss_freq = gen_loss_freq = 5)."""
import collections
import csv
import re
import sys


def load(path):
    lines = open(path).read().splitlines()
    i0 = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines[i0:]):
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        n = re.sub(r"^void ", "", n)[:110]
        agg[n][0] += 1
        agg[n][1] += float(r["Metric Value"]) / 1e3
    return agg


mle, gan = load(sys.argv[1]), load(sys.argv[2])
tm, tg = sum(v[1] for v in mle.values()), sum(v[1] for v in gan.values())
cyc = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for n, v in mle.items():
    cyc[n][0] += 5 * v[0]
    cyc[n][1] += 5 * v[1]
    cyc[n][2] += 5 * v[1]
for n, v in gan.items():
    cyc[n][0] += v[0]
    cyc[n][1] += v[1]
    cyc[n][3] += v[1]
tot = 5 * tm + tg
ours = ("tgan::", "gemm_tc", "relattn", "_kernel<")
print(f"MLE step: {sum(v[0] for v in mle.values())} launches, {tm / 1e3:.2f} ms summed kernel time; "
      f"dis + gen update: {sum(v[0] for v in gan.values())} launches, {tg / 1e3:.2f} ms "
      f"(ncu: serialised, cold caches, host-launched)")
print(f"cycle = 5 x MLE + dis + gen = {tot / 1e3:.2f} ms; MLE share {100 * 5 * tm / tot:.1f} %, adversarial share {100 * tg / tot:.1f} %")
print(f"{'cycle us':>10} {'n':>6} {'share':>6} {'in MLE':>9} {'in GAN':>9}  kernel")
for n, v in sorted(cyc.items(), key=lambda x: -x[1][1])[:45]:
    print(f"{v[1]:10.1f} {v[0]:6d} {100 * v[1] / tot:5.1f}% {v[2]:9.1f} {v[3]:9.1f}  {n}")
lib = sum(v[1] for n, v in cyc.items() if n.startswith("at::") or "cutlass" in n or "cublas" in n or "gemm" in n and "tgan" not in n and "gemm_tc" not in n and "gemm_simt" not in n)
print(f"library (ATen / cuBLAS / CUTLASS) kernels: {100 * lib / tot:.1f} % of the cycle")
