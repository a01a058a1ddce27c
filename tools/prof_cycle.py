"""One iteration-0 of the benchmarked cycle (MLE step + dis update + gen update, experiment_spanbert.yml shapes) between
cudaProfilerStart/Stop, for an ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv \
        python tools/prof_cycle.py [B] [mle,gan] [graphs]

With ``graphs`` the adversarial updates are CUDA-graph replays as in bench.py (ncu profiles the kernel nodes one by one):
that is the path with the side-stream / split single-token backward, which host-launched calls do not take.

The recurrence memory is filled (8 MLE segments) and both adversarial phases have run once before the profiled region.
Kernel shares of the timed cycle = 5 x (MLE step) + dis + gen; tools/cycle_summary.py does that arithmetic."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
what = sys.argv[2] if len(sys.argv) > 2 else "mle,gan"
graphs = len(sys.argv) > 3 and sys.argv[3] == "graphs"
args = types.SimpleNamespace(workload="gan", global_batch=B, scaling="weak", batch_chunk=1, dtype="bf16", kernel_impl=0,
                             no_graphs=not graphs, no_buckets=True)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cyc = bench.Cycle(args, dev, 1, 0)
for _ in range(19 if graphs and "mle" in what else 9):
    cyc.mle_step(False)
for _ in range(3 if graphs else 1):  # graphs: eager warm-up call, capturing call, first replay
    cyc.gan_updates(False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if "mle" in what:
    cyc.mle_step(False)
if "gan" in what:
    cyc.gan_updates(False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled region done")
