"""Helpers shared by the oracle (CPU) and CUDA (GPU) parity tests: load a golden case and replay it."""
import os

import numpy as np
import torch

import txl_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    s = [int(v) for v in z["shape"]]
    shape = O.TxlShape(n_layer=s[0], n_head=s[1], d_model=s[2], d_inner=s[3], n_token=s[4], mem_len=s[5],
                       same_length=bool(s[6]), clamp_len=s[7], pre_lnorm=bool(s[8]))
    return z, shape


def golden_grads(z):
    """{param_name: ('full', array) | ('sample', idx, vals)} and {param_name: norm}."""
    grads, norms = {}, {}
    for k in z.files:
        if k.startswith("gnorm:"):
            norms[k[6:].replace("/", ".")] = float(z[k])
        elif k.startswith("grad:"):
            grads[k[5:].replace("/", ".")] = ("full", z[k])
        elif k.startswith("gidx:"):
            n = k[5:]
            grads[n.replace("/", ".")] = ("sample", z[k], z["gval:" + n])
    return grads, norms


def check_grads(z, got, rtol, atol_frac=1e-3):
    """got: {name: tensor}.  Relative check per tensor: |got-want| <= rtol*|want| + atol_frac*rtol*rms-ish."""
    grads, norms = golden_grads(z)
    worst = 0.0
    for name, spec in grads.items():
        g = got[name].detach().double().cpu()
        n = norms[name]
        assert abs(g.norm().item() - n) <= rtol * max(n, 1e-12) + 1e-12, (name, g.norm().item(), n)
        if spec[0] == "full":
            want = torch.from_numpy(spec[1]).double()
            have = g.reshape(want.shape)
        else:
            want = torch.from_numpy(spec[2]).double()
            have = g.reshape(-1)[torch.from_numpy(spec[1])]
        scale = max(n / max(g.numel(), 1) ** 0.5, 1e-30)  # rms of the tensor
        err = ((have - want).abs() / (want.abs() + scale)).max().item()
        worst = max(worst, err)
        assert err <= rtol, (name, err)
    return worst


def grad_errors(z, got):
    """{name: (frobenius relative error, max elementwise error relative to |want| + rms)} on the golden (sampled)
    gradient entries.  The Frobenius figure is the robust one for low precision: a single ReLU sign flip of a
    near-zero pre-activation changes individual dW entries by one whole term."""
    grads, norms = golden_grads(z)
    out = {}
    for name, spec in grads.items():
        g = got[name].detach().double().cpu()
        n = norms[name]
        if spec[0] == "full":
            want = torch.from_numpy(spec[1]).double()
            have = g.reshape(want.shape)
        else:
            want = torch.from_numpy(spec[2]).double()
            have = g.reshape(-1)[torch.from_numpy(spec[1])]
        scale = max(n / max(g.numel(), 1) ** 0.5, 1e-30)
        frob = ((have - want).norm() / want.norm().clamp_min(1e-30)).item()
        elem = ((have - want).abs() / (want.abs() + scale)).max().item()
        nerr = abs(g.norm().item() - n) / max(n, 1e-30)
        out[name] = (max(frob, nerr), elem)
    return out
