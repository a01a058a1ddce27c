"""2-CTA GEMM check: numerics against torch on the hot shapes + timing against the 1-CTA kernel (env TGAN_GEMM_2CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
from tgan_b200 import lib as L
torch.manual_seed(0)
for name, M, N, K, c32, flags in (("kv_proj", 589824, 1280, 512, False, 0), ("ffn1", 65536, 1024, 512, False, L.EPI_BIAS | L.EPI_RELU),
                                  ("o_proj", 65536, 512, 640, True, L.EPI_ADD_AUX), ("odd", 4100, 776, 520, False, 0)):
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    aux = torch.randn(M, N, device="cuda").bfloat16()
    C = torch.full((M, N), 7.0, device="cuda", dtype=torch.float32 if c32 else torch.bfloat16)
    L.gemm(A, W, C, M=M, N=N, K=K, bias=bias, aux=aux, ldaux=N, flags=flags, impl=2)
    torch.cuda.synchronize()
    rows = torch.cat([torch.arange(0, 300), torch.arange(M // 2 - 150, M // 2 + 150), torch.arange(M - 300, M)]).cuda()
    ref = A[rows].float() @ W.float().t()
    if flags & L.EPI_BIAS: ref = torch.relu(ref + bias)
    if flags & L.EPI_ADD_AUX: ref = ref + aux[rows].float()
    err = (C[rows].float() - ref).abs().max().item()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.gemm(A, W, C, M=M, N=N, K=K, bias=bias, aux=aux, ldaux=N, flags=flags, impl=2); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2] * 1e-3
    print(f"{name:8s} M={M} N={N} K={K} max err {err:.4f}  {t*1e6:8.1f} us  {2.0*M*N*K/t/1e12:7.1f} TFLOP/s  (2CTA={os.environ.get('TGAN_GEMM_2CTA','1')})", flush=True)
