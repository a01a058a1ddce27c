"""GPU: relative-position attention core (forward + backward) against the oracle's closed form."""
import math

import pytest
import torch

import txl_oracle as O

pytestmark = pytest.mark.gpu
HS = 64


def _ref(q, k, v, r, u, vb, reset, B, N, Q, M, dh, mem_len, same_length):
    """fp64 autograd reference built from the oracle pieces.  q [Q,B,N,dh], k/v [K,B,N,dh], r [K,N,dh]."""
    K = M + Q
    mask = O.attn_mask(Q, M, mem_len, same_length, reset, B)
    ac = torch.einsum("ibnd,jbnd->bnij", q + u, k)
    bd = O.rel_shift_gather(torch.einsum("ibnd,jnd->bnij", q + vb, r), Q)
    s = (ac + bd) / math.sqrt(dh)
    s = s.masked_fill(mask[:, None], float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bnij,jbnd->ibnd", p, v), torch.logsumexp(s, -1)


def _pad_heads(x, dtype):  # [..., N, dh] -> [..., N*64]
    out = torch.zeros(*x.shape[:-1], HS, dtype=torch.float32)
    out[..., : x.shape[-1]] = x
    return out.reshape(*x.shape[:-2], -1).to("cuda").to(dtype).contiguous()


CASES = [  # B, N, Q, M, dh, mem_len, same_length, reset
    (2, 3, 5, 0, 10, 8, False, False),
    (3, 2, 8, 16, 10, 16, False, True),
    (2, 2, 8, 12, 50, 12, True, True),
    (2, 10, 16, 24, 50, 24, False, True),
    (4, 2, 1, 37, 50, 64, False, False),      # decode shape
    (1, 2, 128, 128, 50, 128, False, False),  # one full tile
    (2, 2, 128, 256, 50, 256, True, True),
    (2, 1, 64, 200, 50, 200, False, False),
]


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CASES)
def test_relattn_fwd_bwd(case, dtype, impl):
    from tgan_b200 import lib as L
    B, N, Q, M, dh, mem_len, same_length, use_reset = case
    if dtype == torch.float32 and impl == 0:
        pytest.skip("fp32 always runs the SIMT kernels")
    K = M + Q
    g = torch.Generator().manual_seed(B * 1000 + Q * 10 + M)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    q, k, v, r = rnd(Q, B, N, dh), rnd(K, B, N, dh), rnd(K, B, N, dh), rnd(K, N, dh)
    u, vb = 0.3 * rnd(N, dh), 0.3 * rnd(N, dh)
    do = rnd(Q, B, N, dh)
    reset = torch.zeros(B, dtype=torch.bool)
    if use_reset and M > 0:
        reset[B - 1] = True
    # round inputs to the compute dtype first so the reference sees the same operands
    rd = lambda t: t.to(dtype).double()
    q, k, v, r, do = rd(q), rd(k), rd(v), rd(r), rd(do)
    leaves = [t.requires_grad_(True) for t in (q, k, v, r, u, vb)]
    out_ref, lse_ref = _ref(q, k, v, r, u, vb, reset, B, N, Q, M, dh, mem_len, same_length)
    (out_ref * do).sum().backward()

    qd, dod = _pad_heads(q.detach().reshape(Q * B, N, dh), dtype), _pad_heads(do.reshape(Q * B, N, dh), dtype)
    kvd = torch.cat([_pad_heads(k.detach().reshape(K * B, N, dh), dtype),
                     _pad_heads(v.detach().reshape(K * B, N, dh), dtype)], 1).contiguous()
    rdv = _pad_heads(r.detach(), dtype)
    ud, vbd = _pad_heads(u.detach(), torch.float32).view(-1), _pad_heads(vb.detach(), torch.float32).view(-1)
    NH = N * HS
    out = torch.empty(Q * B, NH, device="cuda", dtype=dtype)
    lse = torch.empty(B * N * Q, device="cuda")
    rs = reset.to(torch.uint8).cuda() if use_reset else None
    msl = Q
    if same_length:
        ml = K - mem_len
        msl = Q - ml if ml > 0 else Q
    scale = 1 / math.sqrt(dh)
    L.relattn_fwd(qd, kvd, kvd, 2 * NH, rdv, ud, vbd, rs, out, lse, B, N, Q, M, msl, same_length, scale, 0.0, 0, 0,
                  impl=impl, v_off=NH)
    torch.cuda.synchronize()
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    got = out.float().cpu().view(Q, B, N, HS)
    assert (got[..., :dh].double() - out_ref.detach()).abs().max() < tol
    assert torch.all(got[..., dh:] == 0)
    # the tensor-core path rounds (q + u), (q + v) to bf16 operands: ~0.4% of the score magnitude
    assert (lse.cpu().view(B, N, Q).double() - lse_ref.detach()).abs().max() < (1e-4 if dtype == torch.float32 else 2e-2)

    dq = torch.empty_like(qd)
    dkv = torch.empty_like(kvd)
    dr = torch.empty(K, NH, device="cuda")
    du, dvb = torch.zeros(NH, device="cuda"), torch.zeros(NH, device="cuda")
    delta = torch.empty(B * N * max(Q, M + 1 if Q == 1 else Q), device="cuda")
    L.relattn_bwd(qd, kvd, kvd, 2 * NH, rdv, ud, vbd, rs, out, dod, lse, delta, dq, dkv, dkv, 2 * NH, dr, du, dvb,
                  B, N, Q, M, msl, same_length, scale, 0.0, 0, 0, impl=impl, v_off=NH, dv_off=NH)
    torch.cuda.synchronize()
    gt = 2e-4 if dtype == torch.float32 else 6e-2

    def chk(name, got, want):
        want = want.double()
        err = (got.double().cpu() - want).abs().max().item()
        assert err <= gt * max(1.0, want.abs().max().item()), (name, err, want.abs().max().item())

    chk("dq", dq.float().view(Q, B, N, HS)[..., :dh], leaves[0].grad)
    chk("dk", dkv[:, :NH].float().reshape(K, B, N, HS)[..., :dh], leaves[1].grad)
    chk("dv", dkv[:, NH:].float().reshape(K, B, N, HS)[..., :dh], leaves[2].grad)
    chk("dr", dr.view(K, N, HS)[..., :dh], leaves[3].grad)
    chk("du", du.view(N, HS)[:, :dh], leaves[4].grad)
    chk("dvb", dvb.view(N, HS)[:, :dh], leaves[5].grad)


def test_relattn_dropout_is_consistent_between_fwd_and_bwd():
    """With attention dropout the backward must regenerate the forward's mask: check the directional derivative."""
    from tgan_b200 import lib as L
    B, N, Q, M, dh = 2, 2, 8, 8, 50
    K, NH = M + Q, N * HS
    g = torch.Generator().manual_seed(7)
    mk = lambda rows: _pad_heads(torch.randn(rows, N, dh, generator=g), torch.float32)
    q, kk, vv, r, do = mk(Q * B), mk(K * B), mk(K * B), mk(K), mk(Q * B)
    kv = torch.cat([kk, vv], 1).contiguous()
    u, vb = torch.zeros(NH, device="cuda"), torch.zeros(NH, device="cuda")
    scale = 1 / math.sqrt(dh)

    def fwd(qx):
        out = torch.empty(Q * B, NH, device="cuda")
        lse = torch.empty(B * N * Q, device="cuda")
        L.relattn_fwd(qx, kv, kv, 2 * NH, r, u, vb, None, out, lse, B, N, Q, M, Q, False, scale, 0.3, 42, 17, v_off=NH)
        return out, lse

    out, lse = fwd(q)
    keep_frac = 0.7
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    dr, du, dvb = torch.empty(K, NH, device="cuda"), torch.zeros(NH, device="cuda"), torch.zeros(NH, device="cuda")
    delta = torch.empty(B * N * max(Q, M + 1 if Q == 1 else Q), device="cuda")
    L.relattn_bwd(q, kv, kv, 2 * NH, r, u, vb, None, out, do, lse, delta, dq, dkv, dkv, 2 * NH, dr, du, dvb, B, N, Q, M,
                  Q, False, scale, 0.3, 42, 17, v_off=NH, dv_off=NH)
    eps = 1e-2
    dirn = mk(Q * B)
    o1, _ = fwd(q + eps * dirn)
    o0, _ = fwd(q - eps * dirn)
    fd = ((o1 - o0) * do).sum().item() / (2 * eps)
    an = (dq * dirn).sum().item()
    assert abs(fd - an) < 2e-2 * max(1.0, abs(an)), (fd, an)
    # dropout changed the output (not silently disabled) and is deterministic for a fixed (seed, site)
    out_nodrop = torch.empty_like(out)
    L.relattn_fwd(q, kv, kv, 2 * NH, r, u, vb, None, out_nodrop, lse, B, N, Q, M, Q, False, scale, 0.0, 0, 0, v_off=NH)
    assert (out - out_nodrop).abs().max() > 1e-3
    assert torch.equal(out, fwd(q)[0])


@pytest.mark.parametrize("case", [(2, 2, 128, 256, 256, False, True), (3, 1, 64, 100, 128, True, False),
                                  (2, 3, 128, 1024, 1024, False, False)])
def test_relattn_tcgen05_matches_simt_with_dropout(case):
    """bf16: the tcgen05 kernels (forced) against the SIMT kernels on identical inputs, attention dropout ON --
    both generate the mask from the same counter hash, so outputs and every gradient must agree."""
    from tgan_b200 import lib as L
    B, N, Q, M, mem_len, same_length, use_reset = case
    dh, K, NH = 50, M + Q, N * HS
    g = torch.Generator().manual_seed(Q + M)
    mk = lambda rows: _pad_heads(torch.randn(rows, N, dh, generator=g), torch.bfloat16)
    q, kk, vv, r, do = mk(Q * B), mk(K * B), mk(K * B), mk(K), mk(Q * B)
    kv = torch.cat([kk, vv], 1).contiguous()
    u = _pad_heads(0.3 * torch.randn(N, dh, generator=g), torch.float32).view(-1)
    vb = _pad_heads(0.3 * torch.randn(N, dh, generator=g), torch.float32).view(-1)
    rs = None
    if use_reset:
        rs = torch.zeros(B, dtype=torch.uint8)
        rs[0] = 1
        rs = rs.cuda()
    msl = Q
    if same_length:
        ml = K - mem_len
        msl = Q - ml if ml > 0 else Q
    scale = 1 / math.sqrt(dh)
    res = {}
    for impl in (1, 2):
        out = torch.empty(Q * B, NH, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B * N * Q, device="cuda")
        L.relattn_fwd(q, kv, kv, 2 * NH, r, u, vb, rs, out, lse, B, N, Q, M, msl, same_length, scale, 0.1, 7, 3,
                      impl=impl, v_off=NH)
        dq, dkv = torch.empty_like(q), torch.full_like(kv, 7.0)
        dr = torch.full((K, NH), 7.0, device="cuda")
        du, dvb = torch.zeros(NH, device="cuda"), torch.zeros(NH, device="cuda")
        delta = torch.empty(B * N * max(Q, M + 1 if Q == 1 else Q), device="cuda")
        # both backward passes consume the SAME forward result so only the backward kernels differ
        o_in, l_in = (out, lse) if impl == 1 else (res[1][0], res[1][1])
        L.relattn_bwd(q, kv, kv, 2 * NH, r, u, vb, rs, o_in, do, l_in, delta, dq, dkv, dkv, 2 * NH, dr, du, dvb, B, N,
                      Q, M, msl, same_length, scale, 0.1, 7, 3, impl=impl, v_off=NH, dv_off=NH)
        torch.cuda.synchronize()
        res[impl] = (out, lse, dq.float(), dkv.float(), dr, du, dvb)
    names = ["out", "lse", "dq", "dkv", "dr", "du", "dvb"]
    for nm, a, b in zip(names, res[1], res[2]):
        a, b = a.float(), b.float()
        tol = 3e-2 * max(1.0, a.abs().max().item())
        assert (a - b).abs().max().item() < tol, (nm, (a - b).abs().max().item(), a.abs().max().item())
        rel = (a - b).norm().item() / max(a.norm().item(), 1e-6)
        assert rel < 2e-2, (nm, rel)


def test_relattn_tcgen05_lazy_rescale_large_dynamic_range():
    """Scores whose running maximum keeps growing along the keys (by far more than 2^8 per tile) force the
    forward's lazy rescale of the TMEM-resident output on many tiles; compare with the SIMT kernel."""
    from tgan_b200 import lib as L
    B, N, Q, M, dh = 2, 2, 128, 384, 50
    K, NH = M + Q, N * HS
    g = torch.Generator().manual_seed(5)
    qf = torch.randn(Q * B, N, dh, generator=g)
    kf = torch.randn(K, B, N, dh, generator=g) * torch.linspace(0.1, 12.0, K).view(K, 1, 1, 1)
    vf = torch.randn(K * B, N, dh, generator=g)
    q, kk, vv = _pad_heads(qf, torch.bfloat16), _pad_heads(kf.reshape(K * B, N, dh), torch.bfloat16), _pad_heads(vf, torch.bfloat16)
    r = _pad_heads(torch.randn(K, N, dh, generator=g), torch.bfloat16)
    kv = torch.cat([kk, vv], 1).contiguous()
    u, vb = torch.zeros(NH, device="cuda"), torch.zeros(NH, device="cuda")
    res = {}
    for impl in (1, 2):
        out = torch.empty(Q * B, NH, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B * N * Q, device="cuda")
        L.relattn_fwd(q, kv, kv, 2 * NH, r, u, vb, None, out, lse, B, N, Q, M, Q, False, 1 / math.sqrt(dh), 0.0, 0, 0,
                      impl=impl, v_off=NH)
        torch.cuda.synchronize()
        res[impl] = (out.float(), lse)
    assert torch.isfinite(res[2][0]).all() and torch.isfinite(res[2][1]).all()
    assert (res[1][0] - res[2][0]).abs().max().item() < 3e-2 * max(1.0, res[1][0].abs().max().item())
    assert (res[1][1] - res[2][1]).abs().max().item() < 2e-2 * max(1.0, res[1][1].abs().max().item())
