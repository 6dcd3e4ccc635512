"""Training-step glue for the scripts' hot loop (train_fine.py:173-237,
train_coarse_fineFEAT.py:190-281): the Charades localisation loss, one flat parameter / gradient /
momentum buffer per job, a single NCCL all-reduce of the flat gradient per step, a fused SGD kernel,
and a joint two-stream step (fine stream feeding the coarse stream in memory instead of through
extract_fineFEAT.py's files).

Data parallelism = one process per GPU (torch.distributed, NCCL over NVLink); clips are sharded by
rank, BatchNorm statistics stay per rank exactly as under the reference's nn.DataParallel
(x3d_fine.py:27-29 is a plain BatchNorm3d), and the only collective is the gradient sum."""
import torch
import torch.distributed as dist

import ctypes

from ._lib import call, lib, ptr, stream_ptr

FUSION_LR_MULT = 10.0        # train_coarse_fineFEAT.py:141


class CharadesLossFn(torch.autograd.Function):
    """(cls_loss + loc_loss) * scale with scale = 1/(2*num_steps_per_update) (train_fine.py:199-212,226;
    train_coarse_fineFEAT.py:226-247).  align_corners selects the grid of the F.interpolate to the label length.
    Returns (loss, parts) with parts = [cls_loss, loc_loss] (not differentiable)."""

    @staticmethod
    def forward(ctx, logits, labels, masks, scale, align_corners):
        logits = logits.contiguous().float()
        labels = labels.contiguous().float()
        masks = masks.contiguous().float()
        B, C, T = logits.shape
        TL = labels.shape[2]
        parts = torch.zeros(2, device=logits.device, dtype=torch.float32)
        dlogits = torch.empty_like(logits)
        call("cf_charades_loss", ptr(logits), ptr(labels), ptr(masks), ptr(parts), ptr(dlogits), B, C, T, TL, float(scale),
             int(bool(align_corners)), stream_ptr())
        ctx.save_for_backward(dlogits)
        ctx.mark_non_differentiable(parts)
        return parts.sum() * float(scale), parts

    @staticmethod
    def backward(ctx, dloss, _dparts):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None, None, None, None


def charades_loss(logits, labels, masks, num_steps_per_update=1, align_corners=True):
    """Loss of the scripts' hot loop.  The two scripts resample the logits to the label length on DIFFERENT grids:
    train_fine.py:199 passes align_corners=True (the default here), train_coarse_fineFEAT.py:226 calls
    F.interpolate(per_frame_logits, tl, mode='linear') with PyTorch's default align_corners=False -- use
    ``coarse_charades_loss`` (or align_corners=False) for the coarse script."""
    return CharadesLossFn.apply(logits, labels, masks, 1.0 / (2.0 * num_steps_per_update), align_corners)


def coarse_charades_loss(logits, labels, masks, num_steps_per_update=1):
    """train_coarse_fineFEAT.py:226-247: the same loss on F.interpolate's default (align_corners=False) grid."""
    return charades_loss(logits, labels, masks, num_steps_per_update, align_corners=False)


def is_fusion_param(name):
    """The 10x learning-rate group of train_coarse_fineFEAT.py:137-141."""
    return "rw" in name or "mix" in name


class NativeComm:
    """The C ABI's communicator (cf_comm_*: NCCL resolved with dlopen inside libcfnet_b200.so): ONE in-place fp32 sum
    all-reduce of the flat gradient per step.  The 128-byte NCCL id is created on rank 0 and handed to the other ranks
    through torch.distributed's object broadcast (any initialised backend, gloo is enough) -- the only use of
    torch.distributed on this path."""

    def __init__(self, process_group=None):
        self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        buf = (ctypes.c_ubyte * 128)()
        if self.rank == 0:
            call("cf_comm_unique_id", ctypes.cast(buf, ctypes.c_void_p))
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                                   group=process_group)
        ident = (ctypes.c_ubyte * 128).from_buffer_copy(box[0])
        self._h = ctypes.c_void_p()
        call("cf_comm_init", ctypes.cast(ctypes.byref(self._h), ctypes.c_void_p), self.world, self.rank,
             ctypes.cast(ident, ctypes.c_void_p))

    def allreduce(self, flat):
        call("cf_comm_allreduce", self._h, ptr(flat), flat.numel(), stream_ptr())

    def close(self):
        if self._h.value:
            lib.cf_comm_destroy(self._h)
            self._h = ctypes.c_void_p()


class FlatTrainer:
    """Owns ONE flat fp32 buffer each for parameters, gradients and momentum of a set of modules.

    Every parameter becomes a view of the flat parameter buffer and gets ``_cf_grad``, a view of the
    flat gradient buffer that the weight-gradient kernels accumulate into directly (x3d_ops._flat_grads);
    ``param.grad`` aliases the same view so inspection / checkpointing code keeps working.  ``step()``
    = [one all-reduce of the flat gradient over the data-parallel group] + one fused SGD kernel that
    also re-zeroes the gradient buffer.  Base parameters come first, fusion ('rw'/'mix') parameters
    last, so the two learning-rate groups are two contiguous ranges."""

    def __init__(self, modules, lr, momentum=0.9, weight_decay=1e-5, fusion_lr_mult=FUSION_LR_MULT, process_group=None,
                 native_comm=False):
        """native_comm=True: the gradient all-reduce goes through the C ABI's own NCCL communicator (cf_comm_allreduce)
        instead of torch.distributed.all_reduce (same NCCL, same result)."""
        named = []
        for mi, m in enumerate(modules):
            named += [(f"{mi}.{n}", p) for n, p in m.named_parameters() if p.requires_grad]
        base = [(n, p) for n, p in named if not is_fusion_param(n.split(".", 1)[1])]
        fus = [(n, p) for n, p in named if is_fusion_param(n.split(".", 1)[1])]
        self.names = [n for n, _ in base + fus]
        self.params = [p for _, p in base + fus]
        dev = self.params[0].device
        al = lambda n: (n + 3) // 4 * 4                       # 16-byte aligned segments
        offs, o = [], 0
        for i, p in enumerate(self.params):
            if i == len(base):
                self.n_split = o
            offs.append(o)
            o += al(p.numel())
        if not fus:
            self.n_split = o
        self.n = o
        self.flat_p = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(self.n, device=dev, dtype=torch.float32)
        for p, off in zip(self.params, offs):
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view(p.shape)
            p._cf_grad = self.flat_g[off:off + n].view(p.shape)
            p.grad = p._cf_grad
        self._offs, self._n_base = offs, len(base)
        if dev.type == "cuda":
            from . import x3d_ops
            x3d_ops.PACKS.register(self.flat_p)      # GEMM weights inside this buffer keep persistent packs (x3d_ops.PackCache)
        self.lr, self.momentum, self.weight_decay, self.fusion_lr_mult = lr, momentum, weight_decay, fusion_lr_mult
        self.fusion_lr = None          # explicit learning rate of the fusion group (set by lr_warmup); None = lr * fusion_lr_mult
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.n_params = sum(p.numel() for p in self.params)
        self.comm = NativeComm(process_group) if (native_comm and self.world > 1 and dev.type == "cuda") else None

    def allreduce(self):
        """The one collective of the data-parallel step: sum of the flat gradient over all ranks."""
        if self.flat_g.is_cuda:
            from . import x3d_ops
            x3d_ops.join_side_streams()         # weight-gradient kernels of forked branches write into flat_g untracked
        if self.world > 1:
            if self.comm is not None:
                self.comm.allreduce(self.flat_g)
            else:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)

    def lrs(self):
        """(base-group lr, fusion-group lr) the next update uses."""
        return float(self.lr), float(self.lr * self.fusion_lr_mult if self.fusion_lr is None else self.fusion_lr)

    def sgd(self):
        lr0, lr1 = self.lrs()
        call("cf_sgd_flat", ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_v), self.n, self.n_split, lr0, lr1,
             float(self.momentum), float(self.weight_decay), 1.0 / self.world, stream_ptr())
        if self.flat_p.is_cuda:
            from . import x3d_ops
            x3d_ops.PACKS.weights_changed()          # one launch re-packs every GEMM weight of the step (cf_pw_pack_many)

    def step(self):
        self.allreduce()
        self.sgd()

    def zero_grad(self):
        self.flat_g.zero_()

    # -- checkpoints in torch.optim.SGD's layout: what the scripts save as 'optimizer_state_dict' and load back on resume
    #    (train_fine.py:132-134,245-249; train_coarse_fineFEAT.py:143-145,289-293).  Parameter indices follow the scripts'
    #    group order: base parameters in named_parameters() order, then the 'rw'/'mix' group.
    def _momentum_view(self, i):
        p, off = self.params[i], self._offs[i]
        return self.flat_v[off:off + p.numel()].view(p.shape)

    def state_dict(self):
        lr0, lr1 = self.lrs()
        common = dict(momentum=self.momentum, dampening=0, weight_decay=self.weight_decay, nesterov=False)
        groups = [dict(common, lr=lr0, params=list(range(self._n_base)))]
        if len(self.params) > self._n_base:
            groups.append(dict(common, lr=lr1, params=list(range(self._n_base, len(self.params)))))
        return {"state": {i: {"momentum_buffer": self._momentum_view(i).detach().clone()} for i in range(len(self.params))},
                "param_groups": groups}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        n = sum(len(g["params"]) for g in groups)
        if n != len(self.params):
            raise ValueError(f"optimizer state holds {n} parameters, this trainer {len(self.params)}")
        order = [i for g in groups for i in g["params"]]               # saved index of our i-th parameter
        for i, saved in enumerate(order):
            buf = sd["state"].get(saved, {}).get("momentum_buffer")
            if buf is None:
                self._momentum_view(i).zero_()                          # no step taken yet: torch starts from v = g
            else:
                if tuple(buf.shape) != tuple(self.params[i].shape):
                    raise ValueError(f"momentum buffer {saved}: shape {tuple(buf.shape)} != {tuple(self.params[i].shape)} ({self.names[i]})")
                self._momentum_view(i).copy_(buf)
        self.lr = groups[0]["lr"]
        self.momentum, self.weight_decay = groups[0].get("momentum", self.momentum), groups[0].get("weight_decay", self.weight_decay)
        self.fusion_lr = groups[1]["lr"] if len(groups) > 1 else None


class MultiStepSchedule:
    """optim.lr_scheduler.MultiStepLR(optimizer, milestones, gamma) as the scripts use it (train_fine.py:72,131,256;
    train_coarse_fineFEAT.py: lr_schedule [15,20,25], one step() per epoch) for a FlatTrainer: at every milestone epoch both
    groups' CURRENT learning rates are multiplied by gamma (PyTorch's chainable form), so the fusion group keeps its ratio."""

    def __init__(self, trainer, milestones, gamma=0.1, last_epoch=0):
        self.trainer, self.milestones, self.gamma, self.last_epoch = trainer, sorted(milestones), gamma, last_epoch

    def step(self):
        self.last_epoch += 1
        hits = self.milestones.count(self.last_epoch)
        if hits:
            self.trainer.lr *= self.gamma ** hits
            if self.trainer.fusion_lr is not None:
                self.trainer.fusion_lr *= self.gamma ** hits

    def get_last_lr(self):
        return list(self.trainer.lrs())

    def state_dict(self):
        return {"milestones": list(self.milestones), "gamma": self.gamma, "last_epoch": self.last_epoch,
                "_last_lr": self.get_last_lr()}

    def load_state_dict(self, sd):
        self.milestones, self.gamma, self.last_epoch = sorted(sd["milestones"]), sd["gamma"], sd["last_epoch"]
        if "_last_lr" in sd:
            self.trainer.lr = sd["_last_lr"][0]
            self.trainer.fusion_lr = sd["_last_lr"][1] if len(sd["_last_lr"]) > 1 else None


def lr_warmup(init_lr, cur_steps, warmup_steps, trainer):
    """train_fine.py:258-264 / train_coarse_fineFEAT.py: linear warm-up over the first steps.  As in the reference EVERY
    parameter group is set to lr_scale * init_lr while it is active (the fusion group loses its 10x until the schedule or
    the caller sets it again); outside the window nothing is touched."""
    start_after = 1
    if cur_steps < warmup_steps and cur_steps > start_after:
        lr_scale = min(1., float(cur_steps + 1) / warmup_steps)
        trainer.lr = lr_scale * init_lr
        trainer.fusion_lr = lr_scale * init_lr


def coarse_fine_forward(fine_net, coarse_net, x_fine, start, n_coarse, feat_masks, detach_fine=False, meta=None):
    """Joint two-stream forward.  The fine stream (global_tower=True) runs over the whole clip
    x_fine [B,3,Tf,H,W]; the coarse stream sees the window x_fine[:, :, start:start+n_coarse] and the
    fine features directly from HBM (the reference hands them over through files:
    extract_fineFEAT.py:168-173 -> charades_coarse_fineFEAT.py:84-87, 199-200).  meta = [start,
    n_coarse, Tf, 1] as in charades_coarse_fineFEAT.py:199-200.  detach_fine=True reproduces the
    reference's training semantics (no gradient into the fine stream)."""
    B, _, Tf = x_fine.shape[:3]
    feat, _ = fine_net([x_fine, None])
    if detach_fine:
        feat = {k: v.detach() for k, v in feat.items()}
    if meta is None:                     # pass a prebuilt device tensor when capturing the step in a CUDA graph
        meta = torch.tensor([[float(start), float(n_coarse), float(Tf), 1.0]], device=x_fine.device).repeat(B, 1)
    x_coarse = x_fine[:, :, start:start + n_coarse]
    return coarse_net([x_coarse, feat, feat_masks, 0, meta])


# ----------------------------------------------------------------------------------------
# Evaluation path (train_coarse_fineFEAT.py:213-263)
# ----------------------------------------------------------------------------------------
T_LIM_INFERENCE = 1000       # train_coarse_fineFEAT.py:215


def coarse_forward_chunked(coarse_net, inputs, feat, feat_masks, meta, t_lim=T_LIM_INFERENCE):
    """Validation forward of long videos, train_coarse_fineFEAT.py:215-224: clips longer than t_lim + 5 frames are cut
    into t_lim-frame pieces, each piece sees the whole fine features with its start offset meta[:,0] advanced, and the
    per-frame logits are concatenated along T.  ``meta`` is advanced in place exactly like the reference does."""
    T = inputs.shape[2]
    if T < t_lim + 5:
        return coarse_net([inputs, feat, feat_masks, 0, meta])
    out = []
    for t_ind in range(0, T // t_lim + 1):
        piece = inputs[:, :, t_ind * t_lim:min(T, (t_ind + 1) * t_lim)]
        if piece.shape[2] == 0:                                   # T an exact multiple of t_lim (the reference would fail here)
            break
        out.append(coarse_net([piece, feat, feat_masks, 0, meta]))
        meta[:, 0] += t_lim
    return torch.cat(out, dim=2)


def eval_probs(per_frame_logits, masks, b, n):
    """Multi-view validation scores, train_coarse_fineFEAT.py:231-235: logits [b*n,C,TL] -> max over the n views of
    sigmoid(logits), masked; returns (probs [b,C,TL], logits [b,C,TL])."""
    tl = per_frame_logits.shape[-1]
    lg = per_frame_logits.view(b, n, -1, tl)
    probs = torch.sigmoid(lg).max(dim=1)[0] * masks.unsqueeze(1)
    return probs, lg.max(dim=1)[0]


def localize_samples(probs, labels, valid_t):
    """25 evenly spaced frames of one video (Charades localisation protocol), train_coarse_fineFEAT.py:249-253.
    probs / labels [C,TL] -> ([C,<=25], [C,<=25])."""
    step = int(valid_t / 25.0)
    return probs[:, :valid_t][:, 1::step][:, :25], labels[:, :valid_t][:, 1::step][:, :25]


def charades_csv_rows(name, p1, duration):
    """Rows of the Charades localisation submission file, train_coarse_fineFEAT.py:255-261: (video id, 1 + i*dur/25,
    space-separated class scores) for the <= 25 sampled frames p1 [C,<=25]."""
    a = p1.transpose(0, 1).detach().cpu().numpy()
    return [[name, 1 + i * duration / 25.0, " ".join(str(v) for v in a[i])] for i in range(a.shape[0])]
