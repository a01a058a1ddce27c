"""Times LayerNorm forward / backward at the step's shape (65536 rows x 500) against their byte roofline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from tgan_b200 import lib as L
rows, D, DP = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 500, 512
z = torch.randn(rows, DP, device="cuda"); z[:, D:] = 0
gamma, beta = torch.ones(DP, device="cuda"), torch.zeros(DP, device="cuda")
y = torch.empty(rows, DP, device="cuda", dtype=torch.bfloat16)
mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
dy = torch.randn(rows, DP, device="cuda").bfloat16()
dz, dzd = torch.empty_like(dy), torch.empty_like(dy)
dg, db, ds = (torch.zeros(DP, device="cuda") for _ in range(3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=7):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2] * 1e3
f = t(lambda: L.ln_fwd(z, y, gamma, beta, mean, rstd, rows, D, DP))
b = t(lambda: L.ln_bwd(dy, z, gamma, mean, rstd, dz, dzd, dg, db, rows, D, DP, 0.1, 1, 2, dsum=ds))
bf, bb = rows * DP * (4 + 2), rows * DP * (2 + 4 + 2 + 2)
print(f"ln_fwd {f:7.1f} us  {bf / f / 1e3:6.0f} GB/s   ln_bwd {b:7.1f} us  {bb / b / 1e3:6.0f} GB/s   (rows {rows})")
