"""CPU tests: the oracle (oracle/cf_oracle.py) against the golden vectors produced by the
reference itself (tests/golden/make_golden.py), and -- when /root/reference is present --
against the live reference modules."""
import os

import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from synth import synth_state_dict, synth_tensor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def close(a, b, rtol=1e-4, atol=1e-5):
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= atol + rtol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


def test_interp1d_bit_exact():
    g = load("interp1d")
    ynew, ind = O.interp1d(g["x"], g["y"], g["xnew"])
    assert torch.equal(ind, g["ind"])
    assert torch.equal(ynew, g["ynew"])


@pytest.mark.parametrize("name", ["gridpool_layer", "gridpool_t64"])
def test_gridpool_cdf_and_bins(name):
    g = load(name)
    cdf = O.gridpool_cdf(g["g"])
    close(cdf, g["cdf"], rtol=0, atol=2e-6)
    t = g["x"].shape[2]
    out = O.temporal_lerp(g["x"], g["cdf"])
    close(out, g["out"], rtol=2e-5, atol=1e-6)          # closed form vs 5-D grid_sample (3.5e-6)
    _, i0, w1 = O.sample_coords(g["cdf"], t)
    assert int(i0.min()) >= 0 and int(i0.max()) <= t - 1
    assert bool((i0[:, 1:] >= i0[:, :-1]).all())


def test_gridpool_layer_train_fwd_bwd():
    g = load("gridpool_layer")
    sd = sub(g, "sd_after/")
    # running stats in the fixture are post-update; train-mode forward does not read them
    x = g["x"].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    sd2 = dict(sd)
    sd2.update(params)
    out, cdf = O.gridpool_layer({("pool." + k): v for k, v in sd2.items()}, "pool", x, True)
    close(out, g["out"], rtol=2e-5, atol=1e-6)
    close(cdf, g["cdf"], rtol=0, atol=2e-6)
    ((out * g["gout"]).sum() + (cdf * g["gcdf"]).sum()).backward()
    close(x.grad, g["dx"], rtol=2e-4, atol=1e-6)
    for k, gr in sub(g, "grad/").items():
        # conv{1,2}.bias feed a train-mode BN: their true gradient is 0, both sides are noise
        close(params[k].grad, gr, rtol=5e-4, atol=1e-4 if k in ("conv1.bias", "conv2.bias") else 1e-6)


def test_gather_grads_closed_form():
    g = load("gridpool_layer")
    x = g["x"].clone().requires_grad_(True)
    cdf = g["cdf"].clone().requires_grad_(True)
    (O.temporal_lerp(x, cdf) * g["gout"]).sum().backward()
    close(x.grad, g["gather_dx"], rtol=2e-5, atol=1e-6)
    close(cdf.grad, g["gather_dcdf"], rtol=2e-4, atol=1e-5)


def test_gridunpool():
    g = load("gridunpool")
    x = g["x"].clone().requires_grad_(True)
    cdf = g["cdf"].clone().requires_grad_(True)
    y = O.gridunpool(x, cdf, True)
    close(y, g["y"], rtol=1e-5, atol=1e-6)
    y_up = O.linear_upsample_t(y, (y.shape[2] - 1) * 4)
    close(y_up, g["y_up"], rtol=1e-5, atol=1e-6)
    (y_up * g["gout"]).sum().backward()
    close(x.grad, g["dx"], rtol=1e-5, atol=1e-6)
    close(cdf.grad, g["dcdf"], rtol=2e-4, atol=1e-4)
    yf = O.gridunpool(g["xf"], g["cdf"], False)
    close(yf, g["yf"], rtol=1e-5, atol=1e-6)


def test_gaussian():
    g = load("gaussian")
    cdf = g["cdf"].clone().requires_grad_(True)
    GX = O.gaussian(g["meta"], g["mask"], cdf, int(g["tx"]))
    assert torch.equal(GX, g["GX"]) or (GX - g["GX"]).abs().max() < 1e-7
    (GX * g["gout"]).sum().backward()
    close(cdf.grad, g["dcdf"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,pool,is_mixing", [("rewight", False, True), ("rewight_pool", True, False)])
def test_rewight(name, pool, is_mixing):
    g = load(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in sub(g, "sd/").items()}
    x = g["x"].clone().requires_grad_(True)
    GX = g["GX"].clone().requires_grad_(True)
    bias, scale = O.rewight({"rw." + k: v for k, v in sd.items()}, "rw", x, g["mask"], GX, int(g["height"]), pool, is_mixing)
    close(bias, g["bias"], rtol=1e-5, atol=1e-6)
    close(scale, g["scale"], rtol=1e-5, atol=1e-6)
    ((bias * g["gb"]).sum() + (scale * g["gs"]).sum()).backward()
    close(x.grad, g["dx"], rtol=1e-4, atol=1e-5)
    close(GX.grad, g["dGX"], rtol=1e-4, atol=1e-5)
    for k, gr in sub(g, "grad/").items():
        close(sd[k].grad, gr, rtol=1e-4, atol=1e-5)


def test_mixing():
    g = load("mixing")
    sd = {k: v.clone().requires_grad_(True) for k, v in sub(g, "sd/").items()}
    bs = [g[f"bias{i}"].clone().requires_grad_(True) for i in range(4)]
    ss = [g[f"scale{i}"].clone().requires_grad_(True) for i in range(4)]
    cs, ms = O.mixing({"mix." + k: v for k, v in sd.items()}, "mix", bs, ss, int(g["h"]))
    close(cs, g["cs"], rtol=1e-5, atol=1e-6)
    close(ms, g["ms"], rtol=1e-5, atol=1e-6)
    ((cs * g["gc"]).sum() + (ms * g["gm"]).sum()).backward()
    for i in range(4):
        close(bs[i].grad, g[f"dbias{i}"], rtol=1e-4, atol=1e-5)
        close(ss[i].grad, g[f"dscale{i}"], rtol=1e-4, atol=1e-5)
    for k, gr in sub(g, "grad/").items():
        close(sd[k].grad, gr, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["bottleneck_s2_se", "bottleneck_s1", "bottleneck_s1_se_split2"])
def test_bottleneck(name):
    g = load(name)
    sd_after = sub(g, "sd_after/")
    stride, index, splits = int(g["stride"]), int(g["index"]), int(g["splits"])
    params = {k: v.clone().requires_grad_(True) for k, v in sd_after.items() if v.is_floating_point() and "running" not in k}
    sd = {"blk." + k: v for k, v in {**sd_after, **params}.items()}
    x = g["x"].clone().requires_grad_(True)
    out = O.bottleneck(x, sd, "blk", stride, index, True, splits)
    close(out, g["out"], rtol=2e-5, atol=2e-6)
    (out * g["gout"]).sum().backward()
    close(x.grad, g["dx"], rtol=5e-4, atol=1e-5)
    for k, gr in sub(g, "grad/").items():
        close(params[k].grad, gr, rtol=5e-4, atol=1e-5)
    with torch.no_grad():
        out_eval = O.bottleneck(g["x"], sd, "blk", stride, index, False, splits)
    close(out_eval, g["out_eval"], rtol=2e-5, atol=2e-6)


def _fine_template():
    """Shapes of x3d_fine.generate_model('S', n_classes=10) state dict, rebuilt without the reference."""
    from coarse_fine_networks_b200 import x3d_fine
    return x3d_fine.generate_model("S", n_classes=10, task="loc", base_bn_splits=1, dropout=0.0).state_dict()


def test_fine_net():
    g = load("fine_net")
    sd = synth_state_dict(_fine_template(), 72)
    x = synth_tensor((2, 3, 4, 64, 64), seed=73)
    with torch.no_grad():
        out_eval = O.fine_forward(sd, x, False)
        feats = O.fine_forward(sd, x, False, global_tower=True)
    close(out_eval, g["out_eval"], rtol=1e-4, atol=1e-5)
    for k, v in sub(g, "feat/").items():
        close(feats[k], v, rtol=1e-4, atol=1e-5)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    xg = x.clone().requires_grad_(True)
    out = O.fine_forward({**sd, **params}, xg, True)
    close(out, g["out_train"], rtol=1e-3, atol=1e-4)
    (out * synth_tensor(tuple(out.shape), seed=74)).sum().backward()
    close(xg.grad.sum(dim=(2, 3, 4)), g["dx_sum"], rtol=2e-2, atol=1e-3)
    for k, gr in sub(g, "grad/").items():
        close(params[k].grad, gr, rtol=2e-2, atol=1e-4)


def _coarse_template():
    from coarse_fine_networks_b200 import x3d_coarse
    depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
    m = x3d_coarse.generate_model("M", n_classes=400, feat_depth=depth, task="loc", base_bn_splits=1, dropout=0.0,
                                  t_pool="grid", learnedMixing=True, isMixing=True)
    m.replace_logits(12)
    return m.state_dict()


def test_coarse_net():
    g = load("coarse_net")
    sd = synth_state_dict(_coarse_template(), 82)
    sd["pool_1.conv3.weight"] = sd["pool_1.conv3.weight"] * 8.0
    depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
    B, T, Tf = 1, 8, 12
    x = synth_tensor((B, 3, T, 224, 224), seed=83)
    feat = {k: synth_tensor((B, c, Tf, 7, 7), seed=84 + i).abs() for i, (k, c) in enumerate(depth.items())}
    mask = torch.ones(B, Tf)
    meta = torch.tensor([[2., 8., 12., 1.]])
    with torch.no_grad():
        out_eval = O.coarse_forward(sd, x, feat, mask, meta, False)
    close(out_eval, g["out_eval"], rtol=1e-4, atol=1e-5)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    out = O.coarse_forward({**sd, **params}, x, feat, mask, meta, True)
    close(out, g["out_train"], rtol=1e-3, atol=1e-4)
    (out * synth_tensor(tuple(out.shape), seed=90)).sum().backward()
    # Whole-net gradients at B=1 (train-mode BN over a few hundred positions, ReLU kinks) sit at the
    # fp32 noise floor: the reference's own fp32 gradients (the golden) and this fp32 restatement
    # both differ from an fp64 evaluation of the same graph by 1-2.5 % rel-Linf (SURVEY 8(a)
    # finding 3).  Referee protocol: both must be equally close to the fp64 run.
    cv = lambda t: t.double() if t.is_floating_point() else t
    sd64 = {k: cv(v) for k, v in sd.items()}
    p64 = {k: v.clone().requires_grad_(True) for k, v in sd64.items() if v.is_floating_point() and "running" not in k}
    out64 = O.coarse_forward({**sd64, **p64}, x.double(), {k: v.double() for k, v in feat.items()}, mask.double(),
                             meta.double(), True)
    close(out, out64.float(), rtol=1e-4, atol=1e-5)
    (out64 * synth_tensor(tuple(out.shape), seed=90).double()).sum().backward()
    rl = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    for k, gr in sub(g, "grad/").items():
        e_ref, e_new = rl(gr, p64[k].grad), rl(params[k].grad, p64[k].grad)
        assert e_new <= max(3.0 * e_ref, 1e-4), f"{k}: oracle32 {e_new:.3e} vs reference32 {e_ref:.3e} (both against fp64)"
        cos = torch.nn.functional.cosine_similarity(params[k].grad.double().flatten(), p64[k].grad.flatten(), dim=0).item()
        assert cos >= 0.999, f"{k}: cosine {cos}"
