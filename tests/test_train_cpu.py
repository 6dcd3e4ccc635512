"""CPU tests of the host-side data-parallel logic: flat buffer layout and the world_size-2
gradient all-reduce over gloo (the NCCL path is the same call on CUDA tensors)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _model():
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv1d(3, 5, 1)
            self.rw2 = torch.nn.Linear(5, 3)
            self.mix2 = torch.nn.Linear(3, 2, bias=False)
            self.bn_w = torch.nn.Parameter(torch.ones(7))
    torch.manual_seed(0)
    return M()


def test_flat_layout_and_groups():
    from coarse_fine_networks_b200 import train
    m = _model()
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    tr = train.FlatTrainer([m], lr=0.1)
    assert tr.n % 4 == 0 and tr.n_split % 4 == 0 and 0 < tr.n_split < tr.n
    names = [n.split(".", 1)[1] for n in tr.names]
    first_fusion = min(i for i, n in enumerate(names) if train.is_fusion_param(n))
    assert all(train.is_fusion_param(n) for n in names[first_fusion:])          # fusion group is the tail range
    assert not any(train.is_fusion_param(n) for n in names[:first_fusion])
    lo, hi = tr.flat_p.data_ptr(), tr.flat_p.data_ptr() + tr.n * 4
    for n, p in m.named_parameters():
        assert torch.equal(p.detach(), before[n])                               # values preserved
        assert lo <= p.data_ptr() < hi and p.data_ptr() % 16 == 0               # views of the flat buffer, 16 B aligned
        assert p.grad.data_ptr() == p._cf_grad.data_ptr()
    tr.flat_p.add_(1.0)
    assert torch.allclose(m.conv.weight.detach(), before["conv.weight"] + 1.0)  # aliasing
    assert tr.n_params == sum(p.numel() for p in m.parameters())


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coarse_fine_networks_b200 import train
    m = _model()
    tr = train.FlatTrainer([m], lr=0.1)
    assert tr.world == world
    for i, p in enumerate(tr.params):
        p._cf_grad.fill_(float((rank + 1) * (i + 1)))
    tr.allreduce()                                   # the one collective of the step
    ok = all(torch.all(p._cf_grad == float(sum(r + 1 for r in range(world)) * (i + 1))) for i, p in enumerate(tr.params))
    q.put((rank, bool(ok), float(tr.flat_g.sum())))
    dist.destroy_process_group()


def test_gloo_world2_flat_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2]                     # identical reduced buffers on both ranks


def test_lr_schedule_and_warmup_match_torch_and_the_reference_rule():
    """MultiStepSchedule against torch's MultiStepLR on an SGD with the scripts' two parameter groups (base lr, fusion 10x:
    train_coarse_fineFEAT.py:137-141) with the scripts' milestones [15,20,25]; lr_warmup against the reference's rule
    (train_fine.py:258-264: every group set to lr_scale * init_lr inside the window, untouched outside)."""
    from coarse_fine_networks_b200 import train
    m = _model()
    tr = train.FlatTrainer([m], lr=0.02)
    ref_params = [torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD([{"params": [ref_params[0]]}, {"params": [ref_params[1]], "lr": 0.02 * 10}], lr=0.02, momentum=0.9)
    ref_sched = torch.optim.lr_scheduler.MultiStepLR(opt, [15, 20, 25])
    sched = train.MultiStepSchedule(tr, [15, 20, 25])
    assert tr.lrs() == (0.02, 0.02 * train.FUSION_LR_MULT) and train.FUSION_LR_MULT == 10
    for epoch in range(30):
        opt.step()
        ref_sched.step()
        sched.step()
        want = [g["lr"] for g in opt.param_groups]
        assert all(abs(a - b) <= 1e-12 * b for a, b in zip(tr.lrs(), want)), (epoch, tr.lrs(), want)
    sd = sched.state_dict()
    tr2 = train.FlatTrainer([_model()], lr=0.02)
    s2 = train.MultiStepSchedule(tr2, [1])
    s2.load_state_dict(sd)
    assert s2.last_epoch == 30 and tr2.lrs() == tr.lrs()

    def ref_lr_warmup(init_lr, cur_steps, warmup_steps, o):            # restated from train_fine.py:258-264
        if cur_steps < warmup_steps and cur_steps > 1:
            for pg in o.param_groups:
                pg["lr"] = min(1., float(cur_steps + 1) / warmup_steps) * init_lr
    tr3 = train.FlatTrainer([_model()], lr=0.02)
    opt3 = torch.optim.SGD([{"params": [ref_params[0]]}, {"params": [ref_params[1]], "lr": 0.2}], lr=0.02)
    for step in range(0, 14):
        ref_lr_warmup(0.02, step, 10, opt3)
        train.lr_warmup(0.02, step, 10, tr3)
        assert list(tr3.lrs()) == [g["lr"] for g in opt3.param_groups], step


def test_optimizer_checkpoint_round_trip_with_torch_sgd():
    """FlatTrainer.state_dict() / load_state_dict() speak torch.optim.SGD's layout, so a run resumes from the scripts'
    'optimizer_state_dict' (train_coarse_fineFEAT.py:143-145,289-293) and the scripts can resume from ours."""
    from coarse_fine_networks_b200 import train

    def torch_sgd(model, lr=0.05):
        rw = [p for n, p in model.named_parameters() if "rw" in n or "mix" in n]
        base = [p for n, p in model.named_parameters() if not ("rw" in n or "mix" in n)]
        return torch.optim.SGD([{"params": base}, {"params": rw, "lr": lr * 10}], lr=lr, momentum=0.9, weight_decay=1e-5)

    ref_model = _model()
    opt = torch_sgd(ref_model)
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        for p in ref_model.parameters():
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
    sd = opt.state_dict()

    tr = train.FlatTrainer([_model()], lr=0.01)
    tr.load_state_dict(sd)
    assert tr.lrs() == (0.05, 0.5)
    torch_order = [p for grp in opt.param_groups for p in grp["params"]]
    assert [tuple(p.shape) for p in tr.params] == [tuple(p.shape) for p in torch_order]
    for i, p in enumerate(torch_order):
        assert torch.equal(tr._momentum_view(i), opt.state[p]["momentum_buffer"])
    back = tr.state_dict()
    opt2 = torch_sgd(_model(), lr=0.123)
    opt2.load_state_dict(back)                                           # torch accepts our layout
    assert [grp["lr"] for grp in opt2.param_groups] == [0.05, 0.5]
    for p2, p in zip([q for grp in opt2.param_groups for q in grp["params"]], torch_order):
        assert torch.equal(opt2.state[p2]["momentum_buffer"], opt.state[p]["momentum_buffer"])
    fresh = train.FlatTrainer([_model()], lr=0.01)                       # a state saved before the first step has no buffers
    fresh.flat_v.fill_(7.0)
    fresh.load_state_dict(torch_sgd(_model()).state_dict())
    assert all(float(fresh._momentum_view(i).abs().sum()) == 0.0 for i in range(len(fresh.params)))   # (alignment gaps are not parameters)


def _golden_order():
    import json
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "param_order.json")))


def test_parameter_and_state_dict_order_match_the_reference():
    """named_parameters() order is what torch.optim.SGD checkpoints index by (train_coarse_fineFEAT.py:137-145): it has
    to equal the reference's for fine M / XL and for the coarse net (fixture written by make_golden.py:gen_param_order
    from the unmodified reference)."""
    from coarse_fine_networks_b200 import x3d_coarse, x3d_fine
    gold = _golden_order()
    for v in ("M", "XL"):
        m = x3d_fine.generate_model(v, n_classes=157, task="loc", base_bn_splits=1)
        assert [n for n, _ in m.named_parameters()] == gold[f"fine_{v}"]["params"], v
        assert list(m.state_dict().keys()) == gold[f"fine_{v}"]["state"], v
    depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
    m = x3d_coarse.generate_model("M", n_classes=400, feat_depth=depth, task="loc", base_bn_splits=1, t_pool="grid",
                                  learnedMixing=True, isMixing=True)
    m.replace_logits(157)
    assert [n for n, _ in m.named_parameters()] == gold["coarse_M"]["params"]
    assert list(m.state_dict().keys()) == gold["coarse_M"]["state"]


def test_flat_trainer_indices_line_up_with_the_shipped_optimizer_checkpoints():
    """FlatTrainer's parameter index i (base group in named_parameters() order, then the 'rw'/'mix' group) must carry the
    shape of momentum buffer i of the shipped optimizer_state_dict, for both streams -- resuming maps by position."""
    from coarse_fine_networks_b200 import train, x3d_coarse, x3d_fine
    gold = _golden_order()
    depth = {"layer1": 24, "layer2": 48, "layer3": 96, "layer4": 192, "conv5": 432}
    fine = x3d_fine.generate_model("M", n_classes=157, task="loc", base_bn_splits=1)
    coarse = x3d_coarse.generate_model("M", n_classes=400, feat_depth=depth, task="loc", base_bn_splits=1, t_pool="grid",
                                       learnedMixing=True, isMixing=True)
    coarse.replace_logits(157)
    for tag, net in (("fine", fine), ("coarse", coarse)):
        if f"shipped_{tag}_optimizer" not in gold:
            continue
        o = gold[f"shipped_{tag}_optimizer"]
        tr = train.FlatTrainer([net], lr=0.01)
        assert len(tr.params) == sum(g["n"] for g in o["groups"])
        assert tr._n_base == o["groups"][0]["n"]
        for i, p in enumerate(tr.params):
            assert list(p.shape) == o["shapes"][str(i)], (tag, i, tr.names[i])
