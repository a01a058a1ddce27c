"""CPU: the oracle's LAMB restatement (txl_oracle.lamb_step) against three steps of the UNMODIFIED reference lamb.Lamb
(tests/golden/lamb_tiny.npz, oracle/make_goldens.py::run_lamb_case)."""
import os

import numpy as np
import torch

import txl_oracle as O
from golden_util import GOLD


def lamb_case_tensors(seed):  # same construction as oracle/make_goldens.py::lamb_case_tensors
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(7, 5, generator=g), torch.zeros(4), 20.0 * torch.randn(3000, generator=g),
              0.02 * torch.randn(150, 300, generator=g)]
    grads = [[torch.randn(p.shape, generator=g) * (0.5 + k) for p in params] for k in range(3)]
    return params, grads


def test_oracle_lamb_matches_reference():
    z = np.load(os.path.join(GOLD, "lamb_tiny.npz"))
    params, grads = lamb_case_tensors(int(z["seed"]))
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    for k, gs in enumerate(grads):
        tr = O.lamb_step(params, gs, m, v, float(z["lr"]), weight_decay=float(z["weight_decay"]))
        assert np.allclose(np.array(tr), z[f"trust{k}"], rtol=1e-5)
        for i, p in enumerate(params):
            assert np.allclose(p.numpy(), z[f"p{k}.{i}"], rtol=1e-5, atol=1e-6), (k, i)
    assert z["trust0"][1] == 1.0  # all-zero tensor: trust ratio 1
