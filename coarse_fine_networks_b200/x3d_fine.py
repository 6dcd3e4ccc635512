"""Fine stream (plain X3D) -- drop-in for the reference's ``x3d_fine.py`` module surface.

Same constructor arguments, ``forward([x, masks])`` convention, helper methods and state-dict
key layout as the reference (x3d_fine.py:179-405), so ``train_fine.py`` /
``extract_fineFEAT.py`` call sites and the shipped checkpoints work unchanged.  The nn.Conv3d /
nn.BatchNorm3d / nn.Linear sub-modules are *parameter and buffer holders only*: every forward /
backward runs through the sm_100a kernels of libcfnet_b200.so (x3d_ops.py).  CUDA only.
"""
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import x3d_ops as X


# ----------------------------------------------------------------------------------------
class SubBatchNorm3d(nn.Module):
    """Split batch norm (x3d_fine.py:13-62): ``split_bn`` holds the running statistics of the
    ``num_splits`` sample groups used while training, ``bn`` the aggregated ones used in eval,
    the affine (weight, bias) is kept once.  The normalisation itself is folded into the
    neighbouring conv kernels; calling the module directly runs the standalone kernels."""

    def __init__(self, num_splits, **args):
        super().__init__()
        self.num_splits = num_splits
        self.num_features = args["num_features"]
        self.affine = bool(args.get("affine", True))
        if self.affine:
            self.weight = nn.Parameter(torch.ones(self.num_features))
            self.bias = nn.Parameter(torch.zeros(self.num_features))
        holder = dict(args)
        holder["affine"] = False
        self.bn = nn.BatchNorm3d(**holder)
        holder["num_features"] = self.num_features * self.num_splits
        self.split_bn = nn.BatchNorm3d(**holder)
        self._pending_batches = 0
        self.register_state_dict_pre_hook(lambda module, prefix, keep_vars: module.flush_counters())

    def flush_counters(self):
        """num_batches_tracked is bumped lazily (one host integer per step instead of a kernel)."""
        if self._pending_batches:
            self.split_bn.num_batches_tracked += self._pending_batches
            self._pending_batches = 0

    def aggregate_stats(self):
        """Fold the per-split running statistics into ``bn`` (x3d_fine.py:31-47); call before eval."""
        if not self.split_bn.track_running_stats:
            return
        n = self.num_splits
        means = self.split_bn.running_mean.view(n, -1)
        variances = self.split_bn.running_var.view(n, -1)
        mean = means.mean(0)
        var = variances.mean(0) + ((means - mean) ** 2).mean(0)
        self.bn.running_mean.data = mean.detach()
        self.bn.running_var.data = var.detach()

    def forward(self, x):
        return X.StandaloneBNFn.apply(x, X.BNCfg(self), self.training, self.weight, self.bias)


class Swish(nn.Module):
    """x * sigmoid(x) (x3d_fine.py:65-86); inside Bottleneck it is a conv prologue."""

    def forward(self, x):
        return X.SwishFn.apply(x)


SwishEfficient = X.SwishFn


def conv3x3x3(in_planes, out_planes, stride=1, t_downsample=False):
    """Depthwise 3x3x3 holder (x3d_fine.py:89-97)."""
    st = (stride, stride, stride) if t_downsample else (1, stride, stride)
    return nn.Conv3d(in_planes, out_planes, kernel_size=3, stride=st, padding=1, bias=False, groups=in_planes)


def conv1x1x1(in_planes, out_planes, stride=1, t_downsample=False):
    """Pointwise conv holder (x3d_fine.py:100-105)."""
    st = (stride, stride, stride) if t_downsample else (1, stride, stride)
    return nn.Conv3d(in_planes, out_planes, kernel_size=1, stride=st, bias=False)


def se_width(width, multiplier=0.0625, min_width=8, divisor=8):
    """Bottleneck.round_width (x3d_fine.py:132-143)."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


class _BlockCfg:
    __slots__ = ("stride", "t_stride", "training", "bn1", "bn2", "bn3", "bnd", "arena", "pool", "inference")


class Bottleneck(nn.Module):
    """X3D bottleneck (x3d_fine.py:108-175).  planes = (expanded, out)."""

    def __init__(self, in_planes, planes, stride=1, downsample=None, index=0, base_bn_splits=8, t_downsample=False):
        super().__init__()
        self.index = index
        self.base_bn_splits = base_bn_splits
        mk_bn = lambda c: SubBatchNorm3d(num_splits=base_bn_splits, num_features=c, affine=True)
        self.conv1 = conv1x1x1(in_planes, planes[0])
        self.bn1 = mk_bn(planes[0])
        self.conv2 = conv3x3x3(planes[0], planes[0], stride, t_downsample=t_downsample)
        self.bn2 = mk_bn(planes[0])
        self.conv3 = conv1x1x1(planes[0], planes[1], t_downsample=t_downsample)
        self.bn3 = mk_bn(planes[1])
        self.swish = Swish()
        self.relu = nn.ReLU(inplace=True)
        if index % 2 == 0:
            w = se_width(planes[0])
            self.global_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
            self.fc1 = nn.Conv3d(planes[0], w, kernel_size=1, stride=1)
            self.fc2 = nn.Conv3d(w, planes[0], kernel_size=1, stride=1)
            self.sigmoid = nn.Sigmoid()
        self.downsample = downsample
        self.stride = stride
        self.t_downsample = t_downsample

    round_width = staticmethod(se_width)

    def forward(self, x, pool=None):
        """pool=(rh, rw): also return the (H/rh, W/rw) block average of the output, emitted by the residual-join kernel
        (the global tower's adaptive_avg_pool3d(x, (None,7,7)) of a stage output, x3d_fine.py:345-354) -> (out, pooled)."""
        if self.downsample is not None and not isinstance(self.downsample, nn.Sequential):
            raise NotImplementedError("shortcut_type 'A' (zero-padded identity) is not built; the scripts use 'B'")
        cfg = _BlockCfg()
        cfg.stride, cfg.training = self.stride, self.training
        cfg.t_stride = self.stride if self.t_downsample else 1
        cfg.bn1, cfg.bn2, cfg.bn3 = X.BNCfg(self.bn1), X.BNCfg(self.bn2), X.BNCfg(self.bn3)
        has_se = self.index % 2 == 0
        ds = self.downsample
        cfg.bnd = X.BNCfg(ds[1]) if ds is not None else None
        cfg.arena = getattr(self, "_arena", None)               # set by the owning ResNet for the duration of a forward pass
        cfg.pool = pool
        cfg.inference = (not self.training) and not torch.is_grad_enabled()     # eval without autograd: folded three-launch block
        params = (self.conv1.weight, self.bn1.weight, self.bn1.bias,
                  self.conv2.weight, self.bn2.weight, self.bn2.bias,
                  self.conv3.weight, self.bn3.weight, self.bn3.bias,
                  self.fc1.weight if has_se else None, self.fc1.bias if has_se else None,
                  self.fc2.weight if has_se else None, self.fc2.bias if has_se else None,
                  ds[0].weight if ds is not None else None,
                  ds[1].weight if ds is not None else None, ds[1].bias if ds is not None else None)
        return X.BottleneckFn.apply(x, cfg, *params)


class _SimpleCfg:
    def __init__(self, training, arena=None, **bns):
        self.training = training
        self.arena = arena
        for k, v in bns.items():
            setattr(self, k, X.BNCfg(v))


# ----------------------------------------------------------------------------------------
class ResNet(nn.Module):
    """X3D backbone + per-frame head (x3d_fine.py:179-382)."""

    def __init__(self, block, layers, block_inplanes, n_input_channels=3, conv1_t_size=7, conv1_t_stride=1,
                 shortcut_type='B', widen_factor=1.0, dropout=0.5, n_classes=400, base_bn_splits=8, task='class',
                 extract_feat=False, global_tower=False, t_downsample=False, aux_losses=None):
        super().__init__()
        widths = [tuple(int(c * widen_factor) for c in pair) for pair in block_inplanes]     # (expanded, out) per stage
        stem_c, head_in, head_c = widths[0][1], widths[-1][1], widths[-1][0]
        self.base_bn_splits, self.task, self.t_downsample = base_bn_splits, task, t_downsample
        self.extract_feat, self.global_tower = extract_feat, global_tower
        self.index, self.in_planes = 0, stem_c
        bn = lambda c: SubBatchNorm3d(num_splits=base_bn_splits, num_features=c, affine=True)
        holder = lambda cin, cout, k, s, p, g=1: nn.Conv3d(cin, cout, kernel_size=k, stride=s, padding=p, groups=g, bias=False)

        # parameter holders, registered under the reference's names (state-dict compatibility, x3d_fine.py:210-258)
        self.conv1_s = holder(n_input_channels, stem_c, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        self.conv1_t = holder(stem_c, stem_c, (5, 1, 1), 1, (2, 0, 0), stem_c)
        self.bn1 = bn(stem_c)
        self.relu = nn.ReLU(inplace=True)
        c_in = stem_c
        for stage, (pair, depth) in enumerate(zip(widths, layers), start=1):
            self.add_module(f"layer{stage}", self._make_layer(block, pair, depth, shortcut_type, stride=2, c_in=c_in))
            c_in = pair[1]
        self.in_planes = c_in
        self.conv5 = holder(head_in, head_c, 1, 1, 0)
        self.bn5 = bn(head_c)
        pooled_t = {"class": 1, "loc": None}
        if task in pooled_t:
            self.avgpool = nn.AdaptiveAvgPool3d((pooled_t[task], 1, 1))
        self.fc1 = holder(head_c, 2048, 1, 1, 0)
        self.fc2 = nn.Linear(2048, n_classes)
        self.dropout = nn.Dropout(dropout)
        for conv in (m for m in self.modules() if isinstance(m, nn.Conv3d)):
            nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, block, planes, blocks, shortcut_type, stride=1, c_in=None):
        """One stage (x3d_fine.py:268-300): block 0 carries the stride and the projection shortcut, SE in every even block."""
        c_in = self.in_planes if c_in is None else c_in
        common = dict(base_bn_splits=self.base_bn_splits, t_downsample=self.t_downsample)
        shortcut = None
        if stride != 1 or c_in != planes[1]:
            if shortcut_type == "A":
                shortcut = partial(self._downsample_basic_block, planes=planes[1], stride=stride)
            else:
                shortcut = nn.Sequential(conv1x1x1(c_in, planes[1], stride, t_downsample=self.t_downsample),
                                         SubBatchNorm3d(num_splits=self.base_bn_splits, num_features=planes[1], affine=True))
        stage = [block(c_in if i == 0 else planes[1], planes, stride=stride if i == 0 else 1,
                       downsample=shortcut if i == 0 else None, index=i, **common) for i in range(blocks)]
        self.in_planes = planes[1]
        return nn.Sequential(*stage)

    def _downsample_basic_block(self, x, planes, stride):
        raise NotImplementedError("shortcut_type 'A' is not built (unused by the reference scripts)")

    # -- reference helper surface (x3d_fine.py:309-328)
    def _sub_bns(self):
        return [m for m in self.modules() if isinstance(m, SubBatchNorm3d)]

    def replace_logits(self, n_classes):
        """New classification layer on the device of the old one (x3d_fine.py:309-310)."""
        self.fc2 = nn.Linear(self.fc2.in_features, n_classes).to(self.fc1.weight.device)

    def update_bn_splits_long_cycle(self, long_cycle_bn_scale):
        """Multigrid long cycle: rebuild every split BatchNorm for base_bn_splits * scale splits; returns that count."""
        splits = self.base_bn_splits * long_cycle_bn_scale
        for sbn in self._sub_bns():
            sbn.num_splits = splits
            sbn.split_bn = nn.BatchNorm3d(sbn.num_features * splits, affine=False).to(sbn.weight.device)
        return splits

    def aggregate_sub_bn_stats(self):
        """Fold the split statistics into the eval-mode BatchNorm of every SubBatchNorm3d; returns how many were folded."""
        sbns = self._sub_bns()
        for sbn in sbns:
            sbn.aggregate_stats()
        return len(sbns)

    # -- pieces shared with the coarse stream
    def _open_arena(self, batch, device):
        """One zero-filled fp64 arena for every BatchNorm statistic / backward sum of this pass (x3d_ops.StatsArena)."""
        n = 0
        blocks = [m for m in self.modules() if isinstance(m, Bottleneck)]
        for blk in blocks:
            n += 2 * 4 * batch * max(blk.conv1.weight.shape[0], blk.conv3.weight.shape[0]) * 2
        n += 2 * batch * 2 * (self.conv1_s.weight.shape[0] + self.conv5.weight.shape[0])
        arena = X.StatsArena(n, device)
        for blk in blocks:
            blk._arena = arena
        self._arena = arena
        return arena

    def _close_arena(self):
        for blk in (m for m in self.modules() if isinstance(m, Bottleneck)):
            blk._arena = None
        self._arena = None

    def _stem(self, x):
        return X.StemFn.apply(x, _SimpleCfg(self.training, getattr(self, "_arena", None), bn1=self.bn1), self.conv1_s.weight,
                              self.conv1_t.weight, self.bn1.weight, self.bn1.bias)

    def _conv5_pool(self, x, rh, rw):
        return X.ConvBNReluPoolFn.apply(x, _SimpleCfg(self.training, getattr(self, "_arena", None), bn=self.bn5), self.conv5.weight,
                                        self.bn5.weight, self.bn5.bias, rh, rw)

    def _head(self, pooled):
        """pooled [B,432,T,1,1] (channels-last) -> logits [B,n_classes,T] (x3d_fine.py:370-380)."""
        B, C, T = pooled.shape[:3]
        rows = pooled.permute(0, 2, 3, 4, 1).reshape(B, T, C)
        h = X.LinearRowsFn.apply(rows, self.fc1.weight, None, True)
        h = self.dropout(h)
        out = X.LinearRowsFn.apply(h, self.fc2.weight, self.fc2.bias, False)
        return out.permute(0, 2, 1)

    @staticmethod
    def _pool7(x):
        H, W = x.shape[3], x.shape[4]
        if H % 7 or W % 7:
            raise NotImplementedError("global_tower needs spatial sizes that are multiples of 7 (224x224 clips)")
        return X.AvgPoolFn.apply(x, H // 7, W // 7)

    def forward(self, inp):
        self._open_arena(inp[0].shape[0], inp[0].device)
        try:
            return self._forward(inp)
        finally:
            self._close_arena()

    def _forward(self, inp):
        x, masks = inp
        x = self._stem(x)
        feat_g = {}
        for name in ("layer1", "layer2", "layer3", "layer4"):
            stage = getattr(self, name)
            if not self.global_tower:
                x = stage(x)
                continue
            for blk in list(stage)[:-1]:
                x = blk(x)
            Ho, Wo = (x.shape[3] - 1) // stage[-1].stride + 1, (x.shape[4] - 1) // stage[-1].stride + 1
            if Ho % 7 or Wo % 7:
                raise NotImplementedError("global_tower needs spatial sizes that are multiples of 7 (224x224 clips)")
            x, feat_g[name] = stage[-1](x, pool=(Ho // 7, Wo // 7))     # x3d_fine.py:345-354, fused into the residual join
        B, _, T, H, W = x.shape
        if self.global_tower:
            if H % 7 or W % 7:
                raise NotImplementedError("global_tower needs spatial sizes that are multiples of 7")
            feat_g['conv5'] = self._conv5_pool(x, H // 7, W // 7)   # x3d_fine.py:356-360
            return feat_g, masks
        pooled = self._conv5_pool(x, H, W)                          # [B,432,T,1,1]  AdaptiveAvgPool3d((None,1,1))
        if self.task == 'class':                                    # AdaptiveAvgPool3d((1,1,1)): also over T
            pooled = pooled.mean(dim=2, keepdim=True)
        if self.extract_feat:
            return pooled
        return self._head(pooled)


# X3D variants: per stage (expanded width, output width) and block counts; 'S' and 'M' are the same network (they differ by
# input size only, train_fine.py:59).  x3d_fine.py:388-402
_X3D_M = dict(widths=((54, 24), (108, 48), (216, 96), (432, 192)), depths=(3, 5, 11, 7))
_X3D_XL = dict(widths=((72, 32), (162, 72), (306, 136), (630, 280)), depths=(5, 10, 25, 15))
_VARIANTS = {"S": _X3D_M, "M": _X3D_M, "XL": _X3D_XL}


def get_inplanes(version):
    return [tuple(w) for w in _VARIANTS[version]["widths"]]


def get_blocks(version):
    return list(_VARIANTS[version]["depths"])


def generate_model(x3d_version, **kwargs):
    return ResNet(Bottleneck, get_blocks(x3d_version), get_inplanes(x3d_version), **kwargs)
