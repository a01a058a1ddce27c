"""``nn.Linear`` on the library's tcgen05 GEMM as an autograd function (forward, dgrad, wgrad: three ``tgan_gemm`` calls).

Used by the CNN discriminator ``RelGAN_D`` (model/transformer_gan.py:44-119): its embedding projection, highway and
feature layers are 98 % of its FLOPs (the highway layer alone is a [batch * 64, 1200] x [1200, 1200] product per
call).  bf16 operands, fp32 accumulation; operands whose inner dimension is not a multiple of 8 elements (the 310-token
vocabulary) are zero-padded by ``tgan_convert`` while they are cast.  The backward is NOT itself differentiable: callers
that need a double backward (WGAN-GP on the CNN discriminator) keep the torch path.  CUDA only.
"""
from __future__ import annotations

import torch

from . import lib as L


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _cast_pad(x: torch.Tensor, cols_pad: int) -> torch.Tensor:
    """[rows, cols] (fp32 / bf16, row stride arbitrary) -> contiguous bf16 [rows, cols_pad], pad lanes zero."""
    rows, cols = x.shape
    if x.dtype == torch.bfloat16 and x.stride(1) == 1 and x.stride(0) == cols_pad and cols == cols_pad:
        return x
    if x.stride(1) != 1:
        x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    out = torch.empty(rows, cols_pad, dtype=torch.bfloat16, device=x.device)
    L.convert(x, x.stride(0), out, cols_pad, rows, cols, cols_pad)
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        M, K = x.shape
        N = weight.shape[0]
        Kp = _pad8(K)
        xb, wb = _cast_pad(x.detach(), Kp), _cast_pad(weight.detach(), Kp)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device) if N % 4 == 0 else \
            torch.empty(M, _pad8(N), dtype=torch.float32, device=x.device)[:, :N]
        b32 = None if bias is None else bias.detach().float().contiguous()
        L.gemm(xb, wb, y, M=M, N=N, K=Kp, ldc=y.stride(0), bias=b32, flags=L.EPI_BIAS if b32 is not None else 0,
               impl=L.IMPL_TC)
        ctx.save_for_backward(xb, wb)
        ctx.dims = (M, N, K, Kp, bias is not None, x.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        M, N, K, Kp, has_bias, xdt = ctx.dims
        Np = _pad8(N)
        dyb = _cast_pad(dy, Np)  # [M, Np] bf16 (pad columns zero: they multiply nothing / produce discarded rows)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dx[M, K] = dy[M, N] W[N, K]: B stored [contraction, out] = W as it is
            dxp = torch.empty(M, Kp, dtype=torch.float32, device=dy.device)
            L.gemm(dyb, wb, dxp, transA=False, transB=False, M=M, N=Kp, K=N, lda=Np, ldb=Kp, ldc=Kp, impl=L.IMPL_TC)
            dx = dxp[:, :K].to(xdt) if Kp != K or xdt != torch.float32 else dxp
        if ctx.needs_input_grad[1]:
            # dW[N, K] = dy^T[N, M] x[M, K]: both operands stored [contraction, out]
            dwp = torch.zeros(N, Kp, dtype=torch.float32, device=dy.device)
            L.gemm(dyb, xb, dwp, transA=True, transB=False, M=N, N=Kp, K=M, lda=Np, ldb=Kp, ldc=Kp, flags=L.EPI_ACCUM,
                   impl=L.IMPL_TC)
            dw = dwp[:, :K].contiguous() if Kp != K else dwp
        if has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, dtype=torch.float32, device=dy.device)
            L.colsum(dyb, db, M, N, ld=Np)
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """``F.linear`` for 2-D ``x`` on tgan_gemm (fp32 result)."""
    if not x.is_cuda:
        raise L.TganError("tgan_b200.nn.linear runs on CUDA only (no host fallback)")
    return _LinearFn.apply(x, weight, bias)
