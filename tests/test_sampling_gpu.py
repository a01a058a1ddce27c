"""GPU: batched on-device generation post-processing (tgan_sample_tokens, csrc/sampling.cu) against the oracle's
restatement of generate.py:228-304 (exclude-BOS, empty-bar suppression, temperature, top-k / nucleus / random, one
categorical draw -- torch.multinomial replaced by the inverse CDF of an injected uniform on both sides)."""
import pytest
import torch

import txl_oracle as O

pytestmark = pytest.mark.gpu
V = 310


@pytest.mark.parametrize("technique,temperature,topk,p", [("topk", 0.95, 32, 0.0), ("topk", 1.3, 5, 0.0),
                                                          ("nucleus", 0.9, None, 0.9), ("random", 1.0, None, 0.0),
                                                          ("topk", 0.0, 32, 0.0)])
def test_sampled_ids_and_distribution_match_generate_py(technique, temperature, topk, p):
    from tgan_b200 import lib as L
    g = torch.Generator().manual_seed(17)
    B = 96
    logits = (3.0 * torch.randn(B, 320, generator=g)).float()
    u = torch.rand(B, generator=g)
    suppress = (torch.arange(B) % 3 == 0)
    empty_tok = 101
    ids = torch.empty(B, dtype=torch.int64, device="cuda")
    probs = torch.empty(B, V, device="cuda")
    mode = {"random": 0, "topk": 1, "nucleus": 2}[technique]
    L.sample_tokens(logits.cuda(), ids, V, u=u.cuda(), suppress_empty=suppress.to(torch.uint8).cuda(), probs_out=probs,
                    exclude_bos=True, empty_token=empty_tok, mode=mode, topk=topk or 0, top_p=p, temperature=temperature)
    torch.cuda.synchronize()
    ids, probs = ids.cpu(), probs.cpu()
    checked = 0
    for b in range(B):
        want = O.generation_probs(logits[b, :V].double(), temperature=temperature, technique=technique, topk=topk, p=p,
                                  exclude_bos=True, suppress_empty=bool(suppress[b]), empty_bar_token=empty_tok)
        assert (probs[b].double() - want).abs().max() < 2e-6, (b, (probs[b].double() - want).abs().max())
        assert probs[b, 0] == 0 and (not suppress[b] or probs[b, empty_tok] == 0)
        tok, margin = O.categorical_from_uniform(want, float(u[b]))
        if margin > 1e-5:  # away from a CDF step the draw is bit-exact
            assert int(ids[b]) == tok, (b, int(ids[b]), tok, margin)
            checked += 1
    assert checked > 0.9 * B


def test_device_rng_draws_follow_the_distribution():
    from tgan_b200 import lib as L
    g = torch.Generator().manual_seed(3)
    row = (2.0 * torch.randn(1, 320, generator=g)).float()
    B = 20000
    logits = row.expand(B, 320).contiguous().cuda()
    ids = torch.empty(B, dtype=torch.int64, device="cuda")
    L.sample_tokens(logits, ids, V, mode=1, topk=8, temperature=1.0, seed=5, site=9)
    want = O.generation_probs(row[0, :V].double(), temperature=1.0, technique="topk", topk=8)
    freq = torch.bincount(ids.cpu(), minlength=V).double() / B
    assert (freq - want).abs().max() < 0.015
    assert freq[want == 0].sum() == 0
