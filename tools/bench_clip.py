"""Time the clip input kernel (cf_clip_preprocess) on the bench workload's input: 4 videos x 256 decoded frames
240x320 RGB uint8 -> [4,3,256,224,224] fp32 (train chain: random crop 210 or 168 -> 224, flip; validation chain:
centre crop 240 -> 224).  Algorithmic bytes per frame: 3*crop^2 read + 12*S^2 written.

    python tools/bench_clip.py [T] [once]      # `once`: a single launch (for ncu --set full)
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
from coarse_fine_networks_b200 import spatial_transforms as ST  # noqa: E402

dev = torch.device("cuda")
T = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
once = "once" in sys.argv
B, H, W, S = 4, 240, 320, 224
MEAN, STD = [0.413, 0.368, 0.338], [0.131, 0.125, 0.132]
peak = 6531.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
vids = [torch.randint(0, 256, (T, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(B)]
batch = torch.empty(B, 3, T, S, S, device=dev)
chains = {
    "train (crop+resize+flip)": ST.Compose([ST.MultiScaleRandomCropMultigrid([224 / 256., 224 / 320.], S), ST.RandomHorizontalFlip(),
                                            ST.ToTensor(255), ST.Normalize(MEAN, STD)]),
    "val (centre crop 240->224)": ST.Compose([ST.CenterCropScaled(S), ST.ToTensor(255), ST.Normalize(MEAN, STD)]),
}
for name, tr in chains.items():
    random.seed(0)
    draws = []
    for b in range(B):
        tr.randomize_parameters(S)
        draws.append([(t, {k: v for k, v in vars(t).items() if k in ("scale", "tl_x", "tl_y", "p", "size")}) for t in tr.transforms])

    def run():
        nbytes = 0
        for b in range(B):
            for t, st in draws[b]:
                vars(t).update(st)
            tr.clip(vids[b], out=batch[b])
            nbytes += T * (3 * tr.params(W, H)[2] ** 2 + 12 * S * S)
        return nbytes

    nbytes = run()
    if once:
        torch.cuda.synchronize()
        continue
    run()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()                      # the 4 launches as one graph: the host-side ctypes calls are not timed
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        run()
        with torch.cuda.graph(graph, stream=side):
            run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = sorted(ms)[len(ms) // 2]
    gbs = nbytes / m / 1e6
    print(f"{name:30s} {B}x{T} frames {H}x{W} -> {S}: {m * 1e3:8.1f} us  {gbs:7.1f} GB/s  {100 * gbs / peak:5.1f}% of measured HBM peak "
          f"({B * T / m * 1e3:.0f} frames/s)")
