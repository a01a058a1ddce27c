"""Times the tcgen05 relative-position attention kernels (forward / backward) at experiment_baseline shapes."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from tgan_b200 import lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
N, Q, M, dh, HS = 10, 128, 1024, 50, 64
K, NH = Q + M, N * HS
g = torch.Generator().manual_seed(0)
def mk(rows):
    x = torch.zeros(rows, N, HS)
    x[..., :dh] = torch.randn(rows, N, dh, generator=g)
    return x.reshape(rows, NH).cuda().bfloat16()
q, do, r = mk(Q * B), mk(Q * B), mk(K)
kv = torch.cat([mk(K * B), mk(K * B)], 1).contiguous()
u = torch.zeros(NH, device="cuda"); vb = torch.zeros(NH, device="cuda")
out = torch.empty(Q * B, NH, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B * N * Q, device="cuda")
dq, dkv = torch.empty_like(q), torch.empty_like(kv)
dr = torch.empty(K, NH, device="cuda"); du = torch.zeros(NH, device="cuda"); dvb = torch.zeros(NH, device="cuda")
delta = torch.empty(B * N * Q, device="cuda")
scale = 1 / math.sqrt(dh)
def fwd(): L.relattn_fwd(q, kv, kv, 2 * NH, r, u, vb, None, out, lse, B, N, Q, M, Q, False, scale, drop, 1, 2, impl=2, v_off=NH)
def bwd(): L.relattn_bwd(q, kv, kv, 2 * NH, r, u, vb, None, out, do, lse, delta, dq, dkv, dkv, 2 * NH, dr, du, dvb, B, N, Q, M, Q, False, scale, drop, 1, 2, impl=2, v_off=NH, dv_off=NH)
flops_f = 3 * 2.0 * B * N * Q * K * dh
for name, fn, fl in (("fwd", fwd, flops_f), ("bwd", bwd, flops_f * 8 / 3)):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    print(f"relattn {name} B={B}: {t*1e3:8.3f} ms  {fl/t/1e12:7.1f} algorithmic TFLOP/s (unmasked, d_head 50)", flush=True)
