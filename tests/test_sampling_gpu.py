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


@pytest.mark.parametrize("technique,topk,p,k_empty", [("topk", 32, 0.0, 0), ("topk", 6, 0.0, 2), ("nucleus", None, 0.9, 0)])
def test_batched_generation_loop_matches_generate_py(technique, topk, p, k_empty):
    """MemTransformerLM.generate_batched (device-side loop: single-token forward with the K/V cache + tgan_sample_tokens,
    ids fed back on the device) against generate.py:207-304 restated with the oracle, sequence by sequence, under the
    same injected uniforms: ids bit-exact (fp32 mode) along every sequence up to the first draw whose uniform falls
    within rounding distance of a CDF step."""
    from test_model_gpu import build
    shape = O.TxlShape(n_layer=2, n_head=4, d_model=40, d_inner=72, n_token=310, mem_len=12, same_length=True)
    seed, B, T0, G, temperature = 31, 3, 3, 20, 0.95
    model = build(shape, seed, 1, torch.float32).eval()
    g = torch.Generator().manual_seed(8)
    start = torch.randint(2, 310, (T0, B), generator=g)
    U = torch.rand(G, B, generator=g)
    empty_tok = 37
    if k_empty:  # make the suppressed token likely, so that runs of it (and hence the suppression) actually occur
        with torch.no_grad():
            model.crit.out_layers[0].bias[empty_tok] += 6.0
    ids, _ = model.generate_batched(start.cuda(), G, technique=technique, topk=topk, top_p=p, temperature=temperature,
                                    exclude_bos=True, empty_token=empty_tok, num_empty_tokens_to_ignore=k_empty,
                                    uniforms=U.cuda())
    torch.cuda.synchronize()
    ids = ids.cpu()
    params = {k: v.double() for k, v in O.init_params(shape, seed).items()}
    if k_empty:
        params["crit.out_layers.0.bias"] = params["crit.out_layers.0.bias"].clone()
        params["crit.out_layers.0.bias"][empty_tok] += 6.0
    checked = suppressed = 0
    for b in range(B):  # generate.py handles one sequence at a time
        seq = [int(t) for t in start[:, b]]
        _, mems = O.generate_logits(start[:-1, b:b + 1], None, params, shape)
        for t in range(G):
            lg, mems = O.generate_logits(torch.tensor([[seq[-1]]]), mems, params, shape)
            sup = k_empty > 0 and all(s == empty_tok for s in seq[-k_empty:])
            suppressed += int(sup)
            probs = O.generation_probs(lg[-1, 0], temperature=temperature, technique=technique, topk=topk, p=p,
                                       exclude_bos=True, suppress_empty=sup, empty_bar_token=empty_tok)
            tok, margin = O.categorical_from_uniform(probs, float(U[t, b]))
            got = int(ids[t, b])
            if margin < 2e-4:  # u within the fp32 model error of a CDF step: either neighbour is a legitimate draw
                assert probs[got] > 0, (b, t, got)
            else:
                assert got == tok, (b, t, got, tok, margin)
                checked += 1
            assert got != 0 and not (sup and got == empty_tok)
            seq.append(got)  # follow the device's sequence: later steps stay comparable
    assert checked > 0.85 * B * G
    assert not k_empty or suppressed > 0
