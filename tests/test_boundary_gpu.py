"""GPU: contracts of the drop-in boundary that the reference gets for free from fresh tensors / eager modules and the
ring-buffer memory / packed parameter copies have to enforce explicitly."""
import types

import pytest
import torch

import txl_oracle as O
from test_model_gpu import build

pytestmark = pytest.mark.gpu

SHAPE = O.TxlShape(n_layer=2, n_head=4, d_model=64, d_inner=128, n_token=310, mem_len=32)


def _seg(g, Q, B):
    return torch.randint(2, 310, (Q, B), generator=g).cuda(), torch.randint(2, 310, (Q, B), generator=g).cuda()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_optimizer_that_updates_through_dot_data_is_seen(dtype):
    """lamb.py:116 updates with ``p.data.add_`` -- no autograd version counter moves.  The engine re-packs its
    kernel-private parameter copies on the first forward after a backward, so the update must be visible."""
    g = torch.Generator().manual_seed(0)
    model = build(SHAPE, 3, 16, dtype).train()
    data, target = _seg(g, 16, 4)
    loss0, _ = model(data, target, None, None)
    loss0.mean().backward()
    with torch.no_grad():
        for p in model.parameters():
            p.data.add_(0.05 * torch.sign(p.grad))  # a large ascent step: the loss must move visibly
    loss1, _ = model(data, target, None, None)
    fresh = build(SHAPE, 3, 16, dtype).train()
    fresh.load_state_dict(model.state_dict())
    loss2, _ = fresh(data, target, None, None)
    assert (loss1.mean() - loss0.mean()).abs() > 1e-2, "the parameter update was not picked up"
    assert torch.allclose(loss1, loss2, atol=1e-5 if dtype == torch.float32 else 1e-2)


def test_flat_buffer_optimizer_notifies_the_engine():
    from tgan_b200 import dp
    g = torch.Generator().manual_seed(1)
    model = build(SHAPE, 4, 16, torch.bfloat16).train()
    fp = dp.FlatParams(model.parameters())
    opt = dp.FusedClipAdam(fp, 0.05, clip=1.0)
    data, target = _seg(g, 16, 4)
    with torch.no_grad():
        before, _ = model(data, target, None, None)
    loss, _ = model(data, target, None, None)
    loss.mean().backward()
    opt.step()
    with torch.no_grad():
        after, _ = model(data, target, None, None)  # no backward in between: only notify_params_updated() tells
        opt2_same, _ = model(data, target, None, None)
    assert (after.mean() - before.mean()).abs() > 1e-2
    assert torch.equal(after, opt2_same)


def test_stale_memory_handle_raises_and_reuse_of_an_intact_one_works():
    from tgan_b200 import lib as L
    g = torch.Generator().manual_seed(2)
    model = build(SHAPE, 5, 16, torch.bfloat16).eval()
    with torch.no_grad():
        mems = None
        handles = []
        for s in range(4):
            data, target = _seg(g, 16, 2)
            _, mems = model(data, target, None, mems)
            handles.append(mems)
        data, target = _seg(g, 16, 2)
        # the handle that fed the latest call is intact (its rows were not written): passing it again is legal, as
        # with the reference's fresh tensors, and gives the same result
        a, m_a = model(data, target, None, handles[-2])
        b, m_b = model(data, target, None, handles[-2])
        assert torch.equal(a, b)
        # ... but that second use re-wrote the rows handles[-1] covers, and older handles lost rows long ago
        with pytest.raises(L.TganError, match="stale memory handle"):
            model(data, target, None, handles[-1])
        with pytest.raises(L.TganError, match="stale memory handle"):
            model(data, target, None, handles[0])
        # a materialised copy is an ordinary tensor and can always be passed back
        c, _ = model(data, target, None, m_b.materialize())
        d, _ = model(data, target, None, m_b)
        assert torch.allclose(c, d, atol=2e-2)


@pytest.mark.parametrize("dtype", [torch.float32])
def test_memory_longer_than_mem_len_is_attended_in_full_then_truncated(dtype):
    """reset_length() shrinking mem_len between calls: the reference attends over the whole incoming memory and only
    then keeps the last mem_len rows (mem_transformer.py:556-575, 463-470)."""
    g = torch.Generator().manual_seed(3)
    model = build(SHAPE, 6, 16, dtype).eval()
    p = {k: v.double() for k, v in O.init_params(SHAPE, 6).items()}
    small = O.TxlShape(n_layer=2, n_head=4, d_model=64, d_inner=128, n_token=310, mem_len=8)
    with torch.no_grad():
        mems, mems_o = None, None
        for s in range(3):
            data, target = _seg(g, 16, 2)
            _, mems = model(data, target, None, mems)
            _, mems_o = O.mle_forward(data.cpu(), target.cpu(), None, mems_o, p, SHAPE)
        model.reset_length(16, 8)
        data, target = _seg(g, 16, 2)
        for as_tensor in (False, True):
            loss, new = model(data, target, None, mems.materialize() if as_tensor else mems)
            loss_o, new_o = O.mle_forward(data.cpu(), target.cpu(), None, mems_o, p, small)
            assert (loss.cpu().double() - loss_o).abs().max() < 1e-4, as_tensor
            assert new.size(1) == 8 and (new.materialize().cpu().double() - new_o).abs().max() < 2e-4


def test_pad_type_other_than_model_ignores_reset_mems():
    g = torch.Generator().manual_seed(4)
    model = build(SHAPE, 7, 16, torch.float32).eval()
    with torch.no_grad():
        data, target = _seg(g, 16, 2)
        _, mems = model(data, target, None, None)
        data, target = _seg(g, 16, 2)
        reset = torch.tensor([True, False]).cuda()
        masked, _ = model(data, target, reset, mems)
        plain, _ = model(data, target, None, mems)
        assert (masked[:, 0] - plain[:, 0]).abs().max() > 1e-4 and torch.equal(masked[:, 1], plain[:, 1])
        model.pad_type = "none"  # mem_transformer.py:495-528: the 3-D reset mask exists only for pad_type == 'model'
        unmasked, _ = model(data, target, reset, mems)
        assert torch.equal(unmasked, plain)
