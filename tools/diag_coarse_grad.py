"""Diagnostic (not a test): repeat the whole-coarse-net train-mode gradient comparison of
tests/test_coarse_gpu.py::test_coarse_net_golden several times in one process and print, per
parameter, the error of our gradient against the fp64 oracle (L-inf and L2), so that
run-to-run flips (ReLU kinks / bins at B=1) can be told apart from systematic errors."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synth_state_dict, synth_tensor  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

ge.build()
from coarse_fine_networks_b200 import x3d_coarse as C  # noqa: E402
from oracle import cf_oracle as O  # noqa: E402

depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
d = np.load(os.path.join(ROOT, "tests", "golden", "coarse_net.npz"))
g = {k: torch.from_numpy(d[k]) for k in d.files}
gref = {k[5:]: v for k, v in g.items() if k.startswith("grad/")}


def model():
    m = C.generate_model("M", n_classes=400, feat_depth=depth, task="loc", base_bn_splits=1, dropout=0.0,
                         t_pool="grid", learnedMixing=True, isMixing=True)
    m.replace_logits(12)
    m.rw6.dropout.p = 0.0
    return m


m = model()
sd = synth_state_dict(m.state_dict(), 82)
sd["pool_1.conv3.weight"] = sd["pool_1.conv3.weight"] * 8.0
B, T, Tf = 1, 8, 12
x = synth_tensor((B, 3, T, 224, 224), seed=83)
feat = {k: synth_tensor((B, c, Tf, 7, 7), seed=84 + i).abs() for i, (k, c) in enumerate(depth.items())}
mask = torch.ones(B, Tf)
meta = torch.tensor([[2., 8., 12., 1.]])
gout = synth_tensor((B, 12, 8), seed=90)

cv = lambda t: t.double() if t.is_floating_point() else t
sd64 = {k: cv(v) for k, v in sd.items()}
p64 = {k: v.clone().requires_grad_(True) for k, v in sd64.items() if v.is_floating_point() and "running" not in k}
out64, aux64 = O.coarse_forward({**sd64, **p64}, x.double(), {k: v.double() for k, v in feat.items()}, mask.double(),
                                meta.double(), True, return_aux=True)
(out64 * gout.double()).sum().backward()
print("cdf64", aux64["cdf"].flatten().tolist())

rl = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
l2 = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
prev = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    m = model()
    m.load_state_dict(sd, strict=True)
    m.cuda().train()
    out = m([x.cuda(), {k: v.cuda() for k, v in feat.items()}, mask.cuda(), 0, meta.cuda()])
    (out * gout.cuda()).sum().backward()
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    bad = []
    for k in gref:
        e_ref, e_new = rl(gref[k], p64[k].grad), rl(grads[k], p64[k].grad)
        if not k.startswith("pool_1.") and e_new > max(3 * e_ref, 1e-3):
            bad.append((k, e_new, e_ref, l2(grads[k], p64[k].grad), l2(gref[k], p64[k].grad)))
    print(f"rep {rep}: logits err {rl(out.detach().cpu(), out64.detach()):.2e}; {len(bad)} params over bound")
    for b_ in bad[:12]:
        print("   %-28s linf ours %.3e ref %.3e | l2 ours %.3e ref %.3e" % b_)
    if prev is not None:
        diff = max(rl(grads[k], prev[k].double()) for k in grads if prev[k].abs().max() > 0)
        print(f"   max rel-Linf change of any gradient vs previous rep: {diff:.3e}")
    prev = grads
    if bad:
        k = bad[0][0]
        e = (grads[k].double() - p64[k].grad).abs().flatten()
        top = torch.topk(e, 5)
        print("   worst elements of", k, [(int(i), float(v), float(p64[k].grad.flatten()[i])) for v, i in zip(top.values, top.indices)],
              "max|g64|", float(p64[k].grad.abs().max()))
