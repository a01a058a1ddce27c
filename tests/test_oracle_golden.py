"""CPU: the oracle restatement (oracle/txl_oracle.py) against the golden vectors produced by the unmodified
reference (oracle/make_goldens.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

import golden_util as GU
import txl_oracle as O


@pytest.mark.parametrize("name", ["mle_tiny", "mle_tiny_samelen", "mle_real", "mle_real_q32"])
def test_mle_matches_reference(name):
    z, shape = GU.load(name)
    p = O.init_params(shape, int(z["seed"]), dtype=torch.float64)
    for t in p.values():
        t.requires_grad_(True)
    mems = None
    for s in range(int(z["nseg"])):
        data = torch.from_numpy(z[f"data{s}"])
        target = torch.from_numpy(z[f"target{s}"])
        reset = torch.from_numpy(z[f"reset{s}"])
        loss, mems = O.mle_forward(data, target, reset, mems, p, shape)
        loss.mean().backward()
        np.testing.assert_allclose(loss.detach().numpy(), z[f"loss{s}"], rtol=1e-8, atol=1e-8)
    assert list(mems.shape) == list(z["mems_shape"])
    if "mems" in z.files:
        np.testing.assert_allclose(mems.numpy(), z["mems"], rtol=1e-8, atol=1e-8)
    else:
        np.testing.assert_allclose(mems.numpy()[:, :, :, ::7], z["mems_slice"], rtol=1e-8, atol=1e-8)
    worst = GU.check_grads(z, {k: v.grad for k, v in p.items()}, rtol=1e-7)
    assert worst < 1e-7


def test_generate_and_gumbel_match_reference():
    z, shape = GU.load("generate_tiny")
    p = O.init_params(shape, int(z["seed"]), dtype=torch.float64)
    data = torch.from_numpy(z["data"])
    T = int(z["T"])
    with torch.no_grad():
        full, full_mems = O.generate_logits(data, None, p, shape)
        np.testing.assert_allclose(full.numpy(), z["full_logits"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(full_mems.numpy(), z["full_mems"], rtol=1e-8, atol=1e-8)
        mems, inc = None, []
        for t in range(T):
            lg, mems = O.generate_logits(data[t:t + 1], mems, p, shape)
            inc.append(lg)
        inc = torch.cat(inc, 0)
        np.testing.assert_allclose(inc.numpy(), z["inc_logits"], rtol=1e-8, atol=1e-8)
        # generate.py:309-327 invariance: incremental == one-shot
        np.testing.assert_allclose(inc.numpy(), full.numpy(), rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(mems.numpy(), full_mems.numpy(), rtol=1e-7, atol=1e-8)
        U = torch.from_numpy(z["gumbel_U"])
        _, mems = O.generate_logits(data[:3], None, p, shape)
        inp = data[3:4]
        for t in range(4):
            st, mems, _, ids = O.generate_gumbel(inp, float(z["temperature"]), mems, p, shape, U[t:t + 1])
            np.testing.assert_array_equal(ids.numpy(), z["gumbel_ids"][t:t + 1])
            np.testing.assert_allclose(st.numpy(), z["gumbel_st"][t:t + 1], rtol=1e-8, atol=1e-8)
            inp = ids


def test_mask_closed_form_small_cases():
    # exhaustive small sweep of the closed form against the reference's triu/tril construction
    for Q in (1, 3, 5):
        for M in (0, 2, 7):
            for mem_len in (0, 4, 7, 9):
                K = Q + M
                ones = torch.ones(Q, K)
                ref = torch.triu(ones, 1 + M).bool()
                got = O.attn_mask(Q, M, mem_len, False, None, 1)[0]
                assert torch.equal(ref, got)
                mask_len = K - mem_len
                msl = Q - mask_len if mask_len > 0 else Q
                ref_sl = (torch.triu(ones, 1 + M) + torch.tril(ones, -msl)).bool()
                got_sl = O.attn_mask(Q, M, mem_len, True, None, 1)[0]
                assert torch.equal(ref_sl, got_sl), (Q, M, mem_len)
