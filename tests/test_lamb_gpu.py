"""GPU: tgan_lamb_step through tgan_b200.dp.FusedLamb against three steps of the UNMODIFIED reference lamb.Lamb
(tests/golden/lamb_tiny.npz) and, with gradient clipping / data-parallel scaling, against the oracle restatement."""
import os

import numpy as np
import pytest
import torch

import txl_oracle as O
from golden_util import GOLD
from test_lamb_golden import lamb_case_tensors

pytestmark = pytest.mark.gpu


def test_fused_lamb_matches_reference_golden():
    from tgan_b200 import dp
    z = np.load(os.path.join(GOLD, "lamb_tiny.npz"))
    params, grads = lamb_case_tensors(int(z["seed"]))
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    fp = dp.FlatParams(ps)
    opt = dp.FusedLamb(fp, float(z["lr"]), weight_decay=float(z["weight_decay"]))
    for k, gs in enumerate(grads):
        for p, g in zip(ps, gs):
            p.grad.copy_(g)
        opt.step()
        torch.cuda.synchronize()
        assert np.allclose(opt.trust_ratios().cpu().numpy(), z[f"trust{k}"], rtol=1e-4), k
        for i, p in enumerate(ps):
            assert np.allclose(p.detach().cpu().numpy(), z[f"p{k}.{i}"], rtol=2e-5, atol=2e-6), (k, i)
    assert float(fp.grad.abs().max()) == 0.0  # the step zeroes the flat gradient


def test_fused_lamb_with_clipping_matches_oracle():
    from tgan_b200 import dp
    params, grads = lamb_case_tensors(5)
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    fp = dp.FlatParams(ps)
    clip, lr = 0.25, 0.003
    opt = dp.FusedLamb(fp, lr, clip=clip)
    ref = [p.clone().double() for p in params]
    m = [torch.zeros_like(p) for p in ref]
    v = [torch.zeros_like(p) for p in ref]
    for gs in grads:
        for p, g in zip(ps, gs):
            p.grad.copy_(g)
        opt.step()
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in gs))
        coef = min(1.0, clip / (float(total) + 1e-6))  # clip_grad_norm_ (train.py:914)
        O.lamb_step(ref, [g.double() * coef for g in gs], m, v, lr)
    for p, r in zip(ps, ref):
        assert torch.allclose(p.detach().cpu().double(), r, rtol=2e-5, atol=2e-6)


def test_lamb_optimizer_drop_in_matches_reference_golden_through_zero_grad_and_state_dict():
    """tgan_b200.dp.Lamb used the way train.py uses lamb.Lamb (fresh .grad tensors every step, zero_grad dropping them,
    a state_dict round trip into a new optimizer between steps 2 and 3, an lr scheduler writing param_groups)."""
    from tgan_b200 import dp
    z = np.load(os.path.join(GOLD, "lamb_tiny.npz"))
    params, grads = lamb_case_tensors(int(z["seed"]))
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = dp.Lamb(ps, lr=123.0, weight_decay=float(z["weight_decay"]))
    for k, gs in enumerate(grads):
        if k == 2:  # checkpoint / resume
            sd = opt.state_dict()
            opt = dp.Lamb(ps, lr=123.0, weight_decay=float(z["weight_decay"]))
            opt.load_state_dict(sd)
        opt.param_groups[0]["lr"] = float(z["lr"])  # what a scheduler does
        for p, g in zip(ps, gs):
            p.grad = g.clone().cuda()
        opt.step()
        torch.cuda.synchronize()
        got_tr = np.array([float(opt.state[p]["trust_ratio"]) for p in ps])
        assert np.allclose(got_tr, z[f"trust{k}"], rtol=1e-4), (k, got_tr)
        for i, p in enumerate(ps):
            assert np.allclose(p.detach().cpu().numpy(), z[f"p{k}.{i}"], rtol=2e-5, atol=2e-6), (k, i)
        assert opt.state[ps[0]]["step"] == k + 1
        opt.zero_grad()
    assert set(opt.state[ps[0]]) >= {"step", "exp_avg", "exp_avg_sq", "weight_norm", "adam_norm", "trust_ratio"}


def test_lamb_optimizer_skips_tensors_without_gradient():
    from tgan_b200 import dp
    ps = [torch.nn.Parameter(torch.randn(5, 3).cuda()), torch.nn.Parameter(torch.randn(7).cuda())]
    before = ps[1].detach().clone()
    opt = dp.Lamb(ps, lr=0.01)
    ps[0].grad = torch.randn(5, 3).cuda()
    opt.step()
    assert torch.equal(ps[1].detach(), before) and "exp_avg" not in opt.state[ps[1]]
    ps[1].grad = torch.randn(7).cuda()
    ps[0].grad = torch.randn(5, 3).cuda()
    opt.step()  # the set of trained tensors changed: the flat layout is rebuilt, ps[0]'s moments are carried over
    assert not torch.equal(ps[1].detach(), before) and opt.state[ps[0]]["step"] == 2 and opt.state[ps[1]]["step"] == 1
