"""Times the single-token (Q = 1) attention kernels against their byte stream at the GAN sampling shape.
Usage: python tools/decode_bench.py [B] [M] [reps]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from tgan_b200 import lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
N, Q, dh, HS, NL = 10, 1, 50, 64, 6
K, NH = M + Q, N * HS
g = torch.Generator().manual_seed(0)
mk = lambda r, c: (0.5 * torch.randn(r, c, generator=g)).cuda().bfloat16()
q, do = mk(B, NH), mk(B, NH)
kvs = [mk(K * B, 2 * NH) for _ in range(NL)]  # one K/V cache per layer: > L2 in total
r = mk(K, NH)
u, vb = torch.zeros(NH, device="cuda"), torch.zeros(NH, device="cuda")
out = torch.empty(B, NH, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B * N, device="cuda")
dq, dkv = torch.empty_like(q), torch.empty_like(kvs[0])
dr = torch.empty(K, NH, device="cuda"); du = torch.zeros(NH, device="cuda"); dvb = torch.zeros(NH, device="cuda")
scratch = torch.empty(B * N * K, device="cuda")
scale = 1 / math.sqrt(dh)
def fwd(l): L.relattn_fwd(q, kvs[l], kvs[l], 2 * NH, r, u, vb, None, out, lse, B, N, Q, M, Q, False, scale, 0.1, 1, 2, v_off=NH)
def bwd(l): L.relattn_bwd(q, kvs[l], kvs[l], 2 * NH, r, u, vb, None, out, do, lse, scratch, dq, dkv, dkv, 2 * NH, dr, du, dvb, B, N, Q, M, Q, False, scale, 0.1, 1, 2, v_off=NH, dv_off=NH)
kv_bytes = K * B * 2 * NH * 2
for name, fn, nbytes in (("fwd", fwd, kv_bytes), ("bwd (fused + dR)", bwd, 2 * kv_bytes)):
    for l in range(NL): fn(l)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for l in range(NL): fn(l)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / (reps * NL) * 1e-3
    print(f"decode attention {name} B={B} K={K}: {t*1e6:8.1f} us  {nbytes/t/1e9:7.0f} GB/s of K/V stream ({nbytes/1e6:.0f} MB)", flush=True)
