"""GPU tests of the training-step glue: Charades loss, fused flat SGD, direct flat-gradient
accumulation, the joint two-stream step and its CUDA-graph capture."""
import copy

import pytest
import torch
import torch.nn.functional as F

from synth import synth_state_dict, synth_tensor

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk():
    import __graft_entry__ as ge
    ge.build()
    from coarse_fine_networks_b200 import train, x3d_coarse, x3d_fine
    return type("P", (), dict(T=train, C=x3d_coarse, F=x3d_fine))


def ref_loss(logits, labels, masks, align_corners=True):
    """train_fine.py:199-212,226 (align_corners=True) / train_coarse_fineFEAT.py:226-247 (F.interpolate's default grid)
    restated with the same torch calls."""
    tl = labels.shape[2]
    pl = F.interpolate(logits, tl, mode="linear", align_corners=True) if align_corners else F.interpolate(logits, tl, mode="linear")
    probs = torch.sigmoid(pl) * masks.unsqueeze(1)
    cls = F.binary_cross_entropy(torch.max(probs, dim=2)[0], torch.max(labels, dim=2)[0], reduction="mean")
    loc = F.binary_cross_entropy(probs, labels, reduction="sum") / (torch.sum(masks) * labels.shape[1])
    return (cls + loc) / 2, cls, loc


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("B,C,T,TL", [(2, 7, 16, 160), (3, 157, 64, 640), (1, 5, 8, 8), (2, 9, 64, 37)])
def test_charades_loss_vs_torch(pk, B, C, T, TL, align):
    g = torch.Generator().manual_seed(B * 100 + T)
    logits = (torch.randn(B, C, T, generator=g) * 2).cuda()
    labels = (torch.rand(B, C, TL, generator=g) < 0.05).float().cuda()
    masks = torch.ones(B, TL).cuda()
    masks[0, TL - TL // 4:] = 0
    labels = labels * masks.unsqueeze(1)
    a = logits.clone().requires_grad_(True)
    loss, parts = pk.T.charades_loss(a, labels, masks, align_corners=align)
    (loss * 3.0).backward()
    b = logits.clone().requires_grad_(True)
    rl, rc, rloc = ref_loss(b, labels, masks, align)
    (rl * 3.0).backward()
    if not align:
        l2, _ = pk.T.coarse_charades_loss(logits, labels, masks)
        assert abs(l2.item() - rl.item()) <= 1e-5 * abs(rl.item()) + 1e-7
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item()) + 1e-7
    assert abs(parts[0].item() - rc.item()) <= 1e-5 * abs(rc.item()) + 1e-7
    assert abs(parts[1].item() - rloc.item()) <= 1e-5 * abs(rloc.item()) + 1e-7
    err = (a.grad - b.grad).abs().max().item()
    assert err <= 1e-4 * b.grad.abs().max().item() + 1e-9, err


def test_sgd_flat_vs_torch_sgd(pk):
    torch.manual_seed(0)
    class M(torch.nn.Module):
        def __init__(s):
            super().__init__()
            s.a = torch.nn.Linear(7, 5)
            s.rw2 = torch.nn.Linear(5, 3)           # 'rw' -> fusion group (10x lr)
            s.b = torch.nn.Parameter(torch.randn(11))
    m = M().cuda()
    r = copy.deepcopy(m)
    tr = pk.T.FlatTrainer([m], lr=0.02, momentum=0.9, weight_decay=1e-5)
    base = [p for n, p in r.named_parameters() if "rw" not in n]
    fus = [p for n, p in r.named_parameters() if "rw" in n]
    opt = torch.optim.SGD([{"params": base}, {"params": fus, "lr": 0.2}], lr=0.02, momentum=0.9, weight_decay=1e-5)
    assert tr.n % 4 == 0 and tr.n_split % 4 == 0
    for step in range(4):
        gs = {n: torch.randn_like(p) for n, p in r.named_parameters()}
        for (n, p), (_, q) in zip(m.named_parameters(), r.named_parameters()):
            p._cf_grad.copy_(gs[n])
            q.grad = gs[n].clone()
        tr.step()
        opt.step()
        for (n, p), (_, q) in zip(m.named_parameters(), r.named_parameters()):
            assert (p - q).abs().max().item() <= 1e-6 * (q.abs().max().item() + 1), (step, n)
        assert float(tr.flat_g.abs().max()) == 0.0           # the kernel re-zeroes the gradient buffer


def test_flat_trainer_direct_grads_equal_autograd(pk):
    """Weight-gradient kernels accumulating straight into the flat buffer == grads returned to autograd."""
    m = pk.F.generate_model("S", n_classes=10, task="loc", base_bn_splits=1, dropout=0.0)
    m.load_state_dict(synth_state_dict(m.state_dict(), 72))
    m.cuda().train()
    r = copy.deepcopy(m)
    x = synth_tensor((2, 3, 4, 64, 64), seed=73).cuda()
    go = synth_tensor((2, 10, 4), seed=74).cuda()
    (r([x, None]) * go).sum().backward()
    tr = pk.T.FlatTrainer([m], lr=0.01)
    (m([x, None]) * go).sum().backward()
    rels = []
    for (n, p), (_, q) in zip(m.named_parameters(), r.named_parameters()):
        assert p.grad.data_ptr() == p._cf_grad.data_ptr()
        err = (p._cf_grad - q.grad).abs().max().item()
        # the two runs differ by the order of the fp64 statistics atomics only (a 1-ulp flip of an fp32 BatchNorm table
        # entry), which this small train-mode net (B=2, 64x64, 2x2 positions in layer4) amplifies to the percent level
        # in individual gradients (measured run to run: up to 3 % of the tensor's scale).  A wrong accumulation target
        # would be off by O(1) in every tensor, so: loose per-tensor bound, tight median.
        rel = err / (q.grad.abs().max().item() + 1e-12)
        assert rel <= 0.1, (n, err)
        rels.append(rel)
    rels.sort()
    assert rels[len(rels) // 2] <= 1e-2, rels[len(rels) // 2]


def _joint(pk, n_cls=9):
    depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
    fine = pk.F.generate_model("M", n_classes=n_cls, task="loc", base_bn_splits=1, dropout=0.0, global_tower=True)
    coarse = pk.C.generate_model("M", n_classes=400, feat_depth=depth, task="loc", base_bn_splits=1, dropout=0.0,
                                 t_pool="grid", learnedMixing=True, isMixing=True)
    coarse.replace_logits(n_cls)
    coarse.rw6.dropout.p = 0.0
    fine.load_state_dict(synth_state_dict(fine.state_dict(), 11))
    coarse.load_state_dict(synth_state_dict(coarse.state_dict(), 12))
    return fine.cuda().train(), coarse.cuda().train()


def test_joint_two_stream_step_and_cuda_graph(pk):
    """Fine stream feeds the coarse stream in memory; gradients reach both; a captured CUDA graph of
    fwd + loss + bwd + SGD replays to the same parameters as the eager step."""
    fine, coarse = _joint(pk)
    B, Tf, T, n_cls = 1, 16, 8, 9
    x = synth_tensor((B, 3, Tf, 224, 224), seed=5).cuda()
    labels = (synth_tensor((B, n_cls, 80), seed=6) > 1.2).float().cuda()
    lmask = torch.ones(B, 80).cuda()
    fmask = torch.ones(B, Tf).cuda()
    meta = torch.tensor([[4.0, float(T), float(Tf), 1.0]]).cuda()
    state = (copy.deepcopy(fine.state_dict()), copy.deepcopy(coarse.state_dict()))
    tr = pk.T.FlatTrainer([fine, coarse], lr=0.01)
    assert tr.n_split < tr.n

    def step():
        logits = pk.T.coarse_fine_forward(fine, coarse, x, 4, T, fmask, meta=meta)
        loss, _ = pk.T.charades_loss(logits, labels, lmask)
        loss.backward()
        return logits, loss

    logits, loss = step()
    assert logits.shape == (B, n_cls, (T // 4) * 4) and bool(torch.isfinite(loss))
    gfine = fine.layer2[0].conv1.weight._cf_grad.abs().max().item()
    gcoarse = coarse.layer2[0].conv1.weight._cf_grad.abs().max().item()
    gpool = coarse.pool_1.conv1.weight._cf_grad.abs().max().item()
    assert gfine > 0 and gcoarse > 0 and gpool > 0
    tr.step()
    p_eager = tr.flat_p.clone()
    loss_eager = loss.item()
    # The eager autograd graph must be gone before capture: while it lives, its AccumulateGrad nodes (bound to the
    # stream they were created on) are reused by later forwards and would pull that stream into the capture.
    del logits, loss
    # same step again from the same state, captured in a CUDA graph
    fine.load_state_dict(state[0])
    coarse.load_state_dict(state[1])
    tr.flat_v.zero_()
    tr.flat_g.zero_()
    p0 = tr.flat_p.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):                                   # warm-up on the side stream
            step()
            tr.zero_grad()
    torch.cuda.current_stream().wait_stream(s)
    fine.load_state_dict(state[0])
    coarse.load_state_dict(state[1])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        _, gl = step()
        tr.step()
    tr.flat_p.copy_(p0)
    tr.flat_v.zero_()
    tr.flat_g.zero_()
    fine.load_state_dict(state[0])
    coarse.load_state_dict(state[1])
    g.replay()
    torch.cuda.synchronize()
    assert abs(gl.item() - loss_eager) <= 1e-4 * abs(loss_eager) + 1e-6
    # B=1 train-mode BatchNorm: the gradients themselves carry percent-level fp32 noise between two runs that differ
    # only in atomic ordering (SURVEY 8(a) finding 3), so compare the parameter UPDATES, not bits
    du_e, du_g = (p_eager - p0).double(), (tr.flat_p - p0).double()
    cos = float((du_e * du_g).sum() / (du_e.norm() * du_g.norm()))
    assert cos >= 0.99, cos


def test_fine_stream_alone_in_a_cuda_graph_with_idle_side_streams(pk):
    """The fine stream alone forks only the weight-gradient side stream; the other side streams of the pool (the fusion
    block's) hold no work of the step and must not be pulled into the capture by join_side_streams()."""
    from coarse_fine_networks_b200 import x3d_ops as X
    X.side_streams(torch.device("cuda", torch.cuda.current_device()), 6)        # the full pool exists, as after a joint step
    fine = pk.F.generate_model("S", n_classes=9, task="loc", base_bn_splits=1, dropout=0.0)
    fine.load_state_dict(synth_state_dict(fine.state_dict(), 21))
    fine = fine.cuda().train()
    x = synth_tensor((2, 3, 4, 64, 64), seed=7).cuda()
    labels = (synth_tensor((2, 9, 16), seed=8) > 1.2).float().cuda()
    lmask = torch.ones(2, 16).cuda()
    tr = pk.T.FlatTrainer([fine], lr=0.01)

    def step():
        loss, _ = pk.T.charades_loss(fine([x, None]), labels, lmask)
        loss.backward()
        X.join_side_streams()
        return loss

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            loss_eager = step().item()
            tr.zero_grad()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        gl = step()
    tr.zero_grad()
    g.replay()
    torch.cuda.synchronize()
    assert abs(gl.item() - loss_eager) <= 1e-4 * abs(loss_eager) + 1e-6
    assert float(tr.flat_g.abs().max()) > 0


def test_detached_fine_features_reference_semantics(pk):
    fine, coarse = _joint(pk)
    B, Tf, T = 1, 16, 8
    x = synth_tensor((B, 3, Tf, 224, 224), seed=5).cuda()
    fmask = torch.ones(B, Tf).cuda()
    out = pk.T.coarse_fine_forward(fine, coarse, x, 4, T, fmask, detach_fine=True)
    out.sum().backward()
    assert all(p.grad is None for p in fine.parameters())
    assert coarse.rw2.at1.weight.grad is not None
