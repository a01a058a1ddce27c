"""Per-phase clock64 breakdown of CTA 1 of the tcgen05 GEMM (needs the -DTGAN_PROFILE build).
Usage: TGAN_B200_LIB=transformer-gan_b200/tgan_b200/libtgan_b200_prof.so python tools/gemm_phase_prof.py [B]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from tgan_b200 import lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M, N, K = 1152 * B, 1280, 512
A = torch.randn(M, K, device="cuda").bfloat16()
W = torch.randn(N, K, device="cuda").bfloat16()
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    L.gemm(A, W, C, M=M, N=N, K=K, impl=2)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
L._lib.tgan_debug_gemm_prof.argtypes = [ctypes.c_void_p]
L._lib.tgan_debug_gemm_prof.restype = ctypes.c_int
rc = L._lib.tgan_debug_gemm_prof(buf)
tiles = (M // 128) * (N // 256) / 148.0
print("rc", rc, "tiles per CTA %.1f" % tiles)
for title, base, names in (("TMA producer", 0, ["issue/other", "wait empty"]),
                           ("MMA issuer", 8, ["issue/other", "wait tempty", "wait full"]),
                           ("epilogue warp 2", 16, ["other", "wait tfull", "tmem ld", "wait store-read", "math+smem", "fence+issue"])):
    tot = sum(buf[base:base + 8])
    print(f"{title}: total clk {tot}  ({tot / tiles:.0f} per tile)")
    for i, n in enumerate(names):
        print(f"   {n:16s} {buf[base + i]:10d} {100.0 * buf[base + i] / max(tot, 1):5.1f}%  per tile {buf[base + i] / tiles:7.0f}")
