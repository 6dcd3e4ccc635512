import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cf_oracle as O
from synth import synth_state_dict, synth_tensor
from coarse_fine_networks_b200 import x3d_coarse, x3d_fine
which = sys.argv[1]
torch.set_num_threads(16)
rl = lambda a, b: ((a.detach().cpu().double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
ratio = lambda a, b: ((a.detach().cpu().double() * b).sum() / (b * b).sum().clamp_min(1e-60)).item()
if which == "coarse":
    depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
    m = x3d_coarse.generate_model("M", n_classes=400, feat_depth=depth, task="loc", base_bn_splits=1, dropout=0.0, t_pool="grid", learnedMixing=True, isMixing=True)
    m.replace_logits(12); m.rw6.dropout.p = 0.0
    sd = synth_state_dict(m.state_dict(), 82); sd["pool_1.conv3.weight"] = sd["pool_1.conv3.weight"] * 8.0
    m.load_state_dict(sd); m.cuda().eval()
    B, T, Tf = 1, 8, 12
    x = synth_tensor((B, 3, T, 224, 224), seed=83)
    feat = {k: synth_tensor((B, c, Tf, 7, 7), seed=84 + i).abs() for i, (k, c) in enumerate(depth.items())}
    mask, meta = torch.ones(B, Tf), torch.tensor([[2., 8., 12., 1.]])
    gout = synth_tensor((B, 12, 8), seed=90)
    def oracle(dt):
        cv = lambda t: t.to(dt) if t.is_floating_point() else t
        sdd = {k: cv(v) for k, v in sd.items()}
        ps = {k: v.clone().requires_grad_(True) for k, v in sdd.items() if v.is_floating_point() and "running" not in k}
        o = O.coarse_forward({**sdd, **ps}, cv(x), {k: cv(v) for k, v in feat.items()}, cv(mask), cv(meta), False)
        (o * gout.to(dt)).sum().backward(); return o, ps
    out = m([x.cuda(), {k: v.cuda() for k, v in feat.items()}, mask.cuda(), 0, meta.cuda()])
else:
    m = x3d_fine.generate_model("S", n_classes=10, task="loc", base_bn_splits=1, dropout=0.0)
    sd = synth_state_dict(m.state_dict(), 72); m.load_state_dict(sd); m.cuda().eval()
    x = synth_tensor((2, 3, 4, 64, 64), seed=73); gout = synth_tensor((2, 10, 4), seed=74)
    def oracle(dt):
        cv = lambda t: t.to(dt) if t.is_floating_point() else t
        sdd = {k: cv(v) for k, v in sd.items()}
        ps = {k: v.clone().requires_grad_(True) for k, v in sdd.items() if v.is_floating_point() and "running" not in k}
        o = O.fine_forward({**sdd, **ps}, cv(x), False)
        (o * gout.to(dt)).sum().backward(); return o, ps
    out = m([x.cuda(), None])
o64, p64 = oracle(torch.float64); o32, p32 = oracle(torch.float32)
print("out relerr ours", rl(out, o64), "oracle32", rl(o32, o64))
(out * gout.cuda()).sum().backward()
for k, p in m.named_parameters():
    g64 = p64[k].grad
    if g64 is None or float(g64.abs().max()) == 0: continue
    print(f"{k:40s} ours {rl(p.grad, g64):.3e} ratio {ratio(p.grad, g64):.5f}  oracle32 {rl(p32[k].grad, g64):.3e}")
