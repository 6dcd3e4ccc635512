"""Evaluation path on the GPU: the APMeter kernel against the reference's golden values and the oracle, the
chunked long-video forward against an explicit loop, the multi-view / 25-point helpers against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from synth import synth_tensor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


@pytest.fixture(scope="module")
def pk():
    import __graft_entry__ as ge
    ge.build()
    from coarse_fine_networks_b200 import apmeter, train, x3d_coarse
    return type("P", (), dict(A=apmeter, T=train, C=x3d_coarse))


def test_apmeter_golden(pk):
    g = load("apmeter")
    m = pk.A.APMeter()
    N = g["scores"].shape[0]
    for a in range(0, N, 250):                              # added in pieces, numpy and tensors mixed like the scripts do
        m.add(g["scores"][a:a + 250].numpy(), g["targets"][a:a + 250].numpy())
    ap = m.value()
    assert ap.shape == g["ap"].shape and not ap.is_cuda
    # class 2 holds 100 tied scores: the reference's torch.sort(scores, 0, True) is not stable, so the order inside a tie
    # (and with it the AP, by up to a few percent) is implementation-defined; ours is the stable order
    free = torch.ones(ap.numel(), dtype=torch.bool)
    free[2] = False
    assert torch.allclose(ap[free], g["ap"][free], rtol=0, atol=2e-6), (ap - g["ap"]).abs().max()
    assert abs(float(ap[2] - g["ap"][2])) < 0.05
    mw = pk.A.APMeter()
    for a in range(0, N, 250):
        mw.add(g["scores"][a:a + 250].cuda(), g["targets"][a:a + 250].cuda(), g["weights"][a:a + 250])
    apw = mw.value()
    assert torch.allclose(apw[free], g["ap_weighted"][free], rtol=0, atol=5e-6), (apw - g["ap_weighted"]).abs().max()
    assert abs(float(apw[2] - g["ap_weighted"][2])) < 0.05
    m.reset()
    assert m.value() == 0


def test_apmeter_validation_size_vs_oracle(pk):
    """1 814 validation videos x 25 frames (the Charades localisation setting; 32 of the 157 classes to keep the CPU
    oracle's per-class loop short): 45 chunks of 1024 rows per class, carries across chunks."""
    g = torch.Generator().manual_seed(7)
    N, K = 1814 * 25, 32
    scores = torch.rand(N, K, generator=g)
    targets = (torch.rand(N, K, generator=g) < 0.03).long()
    m = pk.A.APMeter()
    m.add(scores, targets)
    ap = m.value()
    ref = O.average_precision(scores, targets)
    assert torch.allclose(ap, ref, rtol=1e-5, atol=1e-6), (ap - ref).abs().max()


def test_eval_helpers_and_chunked_forward(pk):
    depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
    net = pk.C.generate_model("M", n_classes=400, feat_depth=depth, task="loc", base_bn_splits=1, dropout=0.0, t_pool="grid",
                              learnedMixing=True, isMixing=True)
    net.replace_logits(7)
    net.cuda().eval()
    B, T, Tf = 1, 40, 12
    x = synth_tensor((B, 3, T, 224, 224), seed=3).cuda()
    feat = {k: synth_tensor((B, c, Tf, 7, 7), seed=4 + i).abs().cuda() for i, (k, c) in enumerate(depth.items())}
    fm = torch.ones(B, Tf, device="cuda")
    with torch.no_grad():
        # t_lim = 16: pieces of 16, 16 and 8 frames, start offset advanced by 16 per piece (train_coarse_fineFEAT.py:215-224)
        meta = torch.tensor([[0., 40., 12., 1.]], device="cuda")
        got = pk.T.coarse_forward_chunked(net, x, feat, fm, meta, t_lim=16)
        assert float(meta[0, 0]) == 48.0
        ref, meta2 = [], torch.tensor([[0., 40., 12., 1.]], device="cuda")
        for a in (0, 16, 32):
            ref.append(net([x[:, :, a:a + 16], feat, fm, 0, meta2]))
            meta2[:, 0] += 16
        ref = torch.cat(ref, dim=2)
        assert got.shape == ref.shape == (B, 7, 40) and torch.equal(got, ref)
        short = pk.T.coarse_forward_chunked(net, x[:, :, :16], feat, fm, torch.tensor([[0., 16., 12., 1.]], device="cuda"), t_lim=16)
        assert short.shape == (B, 7, 16)
    # multi-view max + mask, 25-point sampling, CSV rows
    b, n, C, TL = 2, 3, 5, 60
    lg = synth_tensor((b * n, C, TL), seed=9).cuda()
    masks = torch.ones(b, TL, device="cuda")
    masks[1, 50:] = 0
    probs, lmax = pk.T.eval_probs(lg, masks, b, n)
    v = lg.view(b, n, C, TL)
    assert torch.equal(probs, torch.sigmoid(v).max(dim=1)[0] * masks.unsqueeze(1)) and torch.equal(lmax, v.max(dim=1)[0])
    labels = (synth_tensor((C, TL), seed=10) > 1).float().cuda()
    p1, l1 = pk.T.localize_samples(probs[1], labels, 50)
    rp, rl = O.localize_samples(probs[1].cpu(), labels.cpu(), 50)
    assert torch.equal(p1.cpu(), rp) and torch.equal(l1.cpu(), rl) and p1.shape[1] == 25
    rows = pk.T.charades_csv_rows("VID01", p1, 31.0)
    assert len(rows) == 25 and rows[0][0] == "VID01" and rows[2][1] == 1 + 2 * 31.0 / 25.0 and len(rows[0][2].split(" ")) == C
