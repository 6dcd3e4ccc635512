"""One training step of a bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [coarse_fine|fine] [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

ge.build()
which = sys.argv[1] if len(sys.argv) > 1 else "coarse_fine"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else None
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
wl = bench.TrainWorkload(dev, rank=0, which=which, batch=batch)
for _ in range(2):
    wl.fwd_bwd()
    wl.trainer.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
wl.fwd_bwd()
wl.trainer.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one", which, "step, loss", float(wl.loss))
