"""Times the two adversarial updates of bench.py's cycle (graph replay):
    python tools/phase_time.py [side_rows] [B]
side_rows = TxlEngine.side_stream_max_rows (0: everything on one stream)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
args = types.SimpleNamespace(workload="gan", global_batch=B, scaling="weak", batch_chunk=1, dtype="bf16", kernel_impl=0,
                             no_graphs=False, no_buckets=True)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cyc = bench.Cycle(args, dev, 1, 0)
import tgan_b200.engine as E  # noqa: E402
_init = E.TxlEngine.__init__


def patched(self, *a, **k):
    _init(self, *a, **k)
    self.side_stream_max_rows = side


E.TxlEngine.__init__ = patched
cyc.mle_step(False)
L = cyc.L
for phase, opt in (("dis_loss", cyc.dis_opt), ("gen_loss", cyc.gen_opt)):
    def call():
        r = cyc.model(cyc.dev_dis[0], None, None, phase)
        opt.step()
        cyc.fp.zero_grad()
        cyc.dfp.zero_grad()
        return r
    for _ in range(3):
        r = call()
    n0 = L.launch_count()
    ms = bench.time_calls(call, 3)
    print(f"side_rows={side} B={B} {phase}: {ms:8.2f} ms  launches {(L.launch_count() - n0) // 3}  "
          f"value {float(r[phase]):.4f}", flush=True)
