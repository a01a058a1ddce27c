"""Per-phase clock64 breakdown of one CTA of the attention backward (needs the -DTGAN_PROFILE build).
Usage: TGAN_B200_LIB=transformer-gan_b200/tgan_b200/libtgan_b200_prof.so python tools/bwd_phase_prof.py [B]"""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from tgan_b200 import lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
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
L.relattn_fwd(q, kv, kv, 2 * NH, r, u, vb, None, out, lse, B, N, Q, M, Q, False, scale, 0.1, 1, 2, impl=2, v_off=NH)
for _ in range(2):
    L.relattn_bwd(q, kv, kv, 2 * NH, r, u, vb, None, out, do, lse, delta, dq, dkv, dkv, 2 * NH, dr, du, dvb, B, N, Q, M, Q,
                  False, scale, 0.1, 1, 2, impl=2, v_off=NH, dv_off=NH)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
L._lib.tgan_debug_bwd_prof.argtypes = [ctypes.c_void_p]
L._lib.tgan_debug_bwd_prof.restype = ctypes.c_int
rc = L._lib.tgan_debug_bwd_prof(buf)
names = ["prologue", "g_pull", "pairA", "s_wait+ld", "ring+exp", "pairB", "dp_wait+ld", "compute", "(unused)", "wait rdone",
         "wait ks_empty+publish", "epilogue", "-", "-"]
tot = sum(buf[:12])  # 12, 13 are sub-intervals of flush_keys / flush_dr
print("rc", rc, "total clk", tot, "(warp 0 lane 0 of CTA 200; 18 tiles)")
for n, v in zip(names, buf):
    print(f"{n:12s} {v:9d} {100.0*v/tot:5.1f}%   per tile {v/18:8.0f}")

mn = ["issue/other", "k_full", "s_empty", "v_full", "dp_empty", "rg_full", "g_empty", "p_full(+kd_empty)", "rd_full", "dr_empty"]
tot2 = sum(buf[16:26])
print("MMA thread: total clk", tot2)
for n, v in zip(mn, buf[16:26]):
    print(f"{n:12s} {v:9d} {100.0*v/tot2:5.1f}%   per tile {v/18:8.0f}")
