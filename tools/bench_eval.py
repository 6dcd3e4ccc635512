"""Eval-mode forward of the fine global tower ([B,3,256,224,224], the feature-extraction / validation path of
extract_fineFEAT.py:137-173) with the folded three-launch blocks (torch.no_grad) against the training launch sequence with
running statistics (the same modules called with autograd enabled): clips/s, launches of our kernels per forward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
from coarse_fine_networks_b200 import _lib, x3d_fine  # noqa: E402
from synth import synth_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = x3d_fine.generate_model("M", n_classes=157, task="loc", base_bn_splits=1, dropout=0.0, global_tower=True)
net.load_state_dict(synth_state_dict(net.state_dict(), 1))
net.cuda().eval()
x = torch.randn(B, 3, 256, 224, 224, device="cuda")


def run(folded):
    ctx = torch.no_grad() if folded else torch.enable_grad()
    with ctx:
        feat, _ = net([x, None])
    return feat


outs = {}
for folded in (False, True):
    for _ in range(2):
        outs[folded] = run(folded)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(folded)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{'folded (no_grad)' if folded else 'training sequence'}: {ms:8.2f} ms per forward = {B / ms * 1e3:7.1f} clips/s, "
          f"{(_lib.launch_count() - n0) // 5} launches")
for k in outs[True]:
    a, b = outs[True][k], outs[False][k].detach()
    print(k, "rel diff folded vs sequence", float((a - b).abs().max() / b.abs().max()))
