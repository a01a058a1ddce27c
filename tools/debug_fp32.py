"""Debug helper: fp32 engine vs fp64 oracle at real model size; prints per-parameter gradient errors."""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "transformer-gan_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import txl_oracle as O
from test_model_gpu import build

shape = O.TxlShape(n_layer=int(sys.argv[1]) if len(sys.argv) > 1 else 2, n_head=10, d_model=500, d_inner=1000, n_token=310, mem_len=24)
Q, B, nseg = 16, 2, int(sys.argv[2]) if len(sys.argv) > 2 else 1
model = build(shape, 13, Q, torch.float32).train()
p = {k: v.double().requires_grad_(True) for k, v in O.init_params(shape, 13).items()}
g = torch.Generator().manual_seed(0)
mems = mo = None
for s in range(nseg):
    data = torch.randint(2, 310, (Q, B), generator=g); tgt = torch.randint(2, 310, (Q, B), generator=g)
    reset = torch.zeros(B, dtype=torch.bool)
    if s == 2 and len(sys.argv) > 3: reset[1] = True
    loss, mems = model(data.cuda(), tgt.cuda(), reset.cuda(), mems)
    loss.mean().backward()
    lo, mo = O.mle_forward(data, tgt, reset, mo, p, shape)
    lo.mean().backward()
    print("seg", s, "loss err", (loss.detach().cpu().double() - lo.detach()).abs().max().item())
    mm = mems.materialize().cpu().double()
    for l in range(shape.n_layer + 1):
        print("  slab", l, "err", (mm[l] - mo[l]).abs().max().item(), "mag", mo[l].abs().max().item())
for k, v in model.named_parameters():
    w = p[k].grad
    e = (v.grad.cpu().double() - w)
    print(f"{k:45s} frob {e.norm().item() / w.norm().item():.2e}  max {e.abs().max().item():.2e} / {w.abs().max().item():.2e}")
