"""Times the tcgen05 GEMM on the hot shapes of the experiment_baseline step (CUDA events, L2-cold operands)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from tgan_b200 import lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
R, KR = 128 * B, 1152 * B
shapes = [  # name, M, N, K, transA, transB, c fp32?
    ("kv_proj", KR, 1280, 512, False, True, False),
    ("q_proj", R, 640, 512, False, True, False),
    ("o_proj", R, 512, 640, False, True, True),
    ("ffn1", R, 1024, 512, False, True, False),
    ("ffn2", R, 512, 1024, False, True, True),
    ("logits", R, 310, 512, False, True, True),
    ("wgrad_kv", 1280, 512, KR, True, False, True),
    ("wgrad_ffn1", 1024, 512, R, True, False, True),
]
for name, M, N, K, tA, tB, c32 in shapes:
    A = torch.randn((K, M) if tA else (M, K), device="cuda").bfloat16()
    Bm = torch.randn((N, K) if tB else (K, N), device="cuda").bfloat16()
    C = torch.empty(M, N, device="cuda", dtype=torch.float32 if c32 else torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        L.gemm(A, Bm, C, transA=tA, transB=tB, M=M, N=N, K=K, impl=2)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.gemm(A, Bm, C, transA=tA, transB=tB, M=M, N=N, K=K, impl=2)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2] * 1e-3
    print(f"{name:12s} M={M:7d} N={N:5d} K={K:7d}  {t*1e6:9.1f} us  {2.0*M*N*K/t/1e12:7.1f} TFLOP/s", flush=True)
    del A, Bm, C
