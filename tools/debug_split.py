"""Per-parameter difference between the eager and the graph-captured generator update on the tiny GAN fixture."""
import os, sys, tempfile, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "transformer-gan_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import test_gan_gpu as T
model, z, shape = T._build_gan("gan_bert_tiny", pathlib.Path(tempfile.mkdtemp()), torch.bfloat16)
data = torch.from_numpy(z["data"]).cuda()
U = torch.from_numpy(z["U"]).cuda()
model.gumbel_noise_source = lambda step, shp: U[step:step + 1]
model.sources_graph_safe = True
def run():
    model.zero_grad(set_to_none=False)
    model(data, None, None, "gen_loss")
    torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in model.generator.named_parameters() if p.grad is not None}
for prm in model.parameters():
    if prm.grad is None and prm.requires_grad:
        prm.grad = torch.zeros_like(prm)
model.use_cuda_graphs = False
ge = run()
model.use_cuda_graphs = True
for rep in range(3):
    gg = run()
    bad = [(k, float((gg[k] - ge[k]).norm() / (ge[k].norm() + 1e-12))) for k in ge]
    bad = [b for b in bad if b[1] > 5e-3]
    print("rep", rep, "params off by > 5e-3:", bad[:12])
k = "layers.0.dec_attn.qkv_net.weight"
d = (gg[k] - ge[k]).view(3, shape.n_head, -1, gg[k].shape[1]).norm(dim=(2, 3))
print("qkv_net.weight error by (q|k|v, head):", d, "norms", ge[k].view(3, shape.n_head, -1, gg[k].shape[1]).norm(dim=(2, 3)))
