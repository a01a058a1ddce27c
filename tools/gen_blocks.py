"""Diagnosis: decode ms/step in blocks of 10 steps across the ring re-layout (B=128, memory 4146)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch
import bench as BN
import mem_transformer as MT
dev = torch.device("cuda", 0)
model = MT.MemTransformerLM(BN.make_cfg(), 310, 0)
BN.init_like_train_py(model, 1111)
model = model.to(dev).eval()
B, mem_len = 128, 4146
g = torch.Generator().manual_seed(3)
with torch.no_grad():
    mems = None
    model.reset_length(128, mem_len)
    for _ in range(33):
        _, mems = model.forward_generate(torch.randint(2, 310, (128, B), generator=g).to(dev), mems)
    model.reset_length(1, mem_len)
    tok = torch.randint(2, 310, (1, B), generator=g).to(dev)
    for blk in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            _, mems = model.forward_generate(tok, mems)
        torch.cuda.synchronize()
        print(f"steps {blk*10:3d}-{blk*10+9:3d}: {1e2*(time.perf_counter()-t0):.2f} ms/step  cap {mems.capacity} start {mems.start} kv {mems.kv['lo']}-{mems.kv['hi']} tag {mems.kv['tag']} epoch {model._engine.pack_epoch}", flush=True)
