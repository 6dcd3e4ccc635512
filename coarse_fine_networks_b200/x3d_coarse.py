"""Coarse stream (X3D + Grid Pool + Multi-stage Fusion + Grid Unpool) -- drop-in for the
reference's ``x3d_coarse.py`` module surface (x3d_coarse.py:175-750).

Same constructor arguments, ``forward([x, feat, feat_masks, i, meta])`` convention, helper
methods and state-dict key layout as the reference, so ``train_coarse_fineFEAT.py`` call sites
and the shipped ``coarse_fineFEAT_charades_*.pt`` checkpoint work unchanged.  nn.Conv1d / nn.Conv3d
/ nn.Linear sub-modules are parameter holders only; all arithmetic runs in libcfnet_b200.so.

What differs from the reference on purpose (same results):
  * the whole fusion block is evaluated at the 7x7 resolution of the fine features; the
    reference's up-sampled 6-D broadcast products are never built;
  * Grid Pool sampling is a temporal lerp gather (no meshgrid, no 5-D grid tensor);
  * activations are channels-last in HBM; the module boundary accepts and returns the
    reference's logical NCTHW shapes.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fusion_ops as FU
from . import gridpool_ops as G
from . import x3d_ops as X
from .interp1d import Interp1d  # noqa: F401  (re-exported like the reference's import)
from .x3d_fine import (Bottleneck, SubBatchNorm3d, Swish, SwishEfficient, _SimpleCfg, conv1x1x1,  # noqa: F401
                       conv3x3x3, get_blocks, get_inplanes)
from .x3d_fine import ResNet as _FineResNet


# ----------------------------------------------------------------------------------------
class RewightLayer(nn.Module):
    """Self-attention filter + Gaussian-aligned aggregation + the two k=1 MLPs that emit the
    shift ("bias") and scale maps (x3d_coarse.py:175-247)."""

    def __init__(self, channels, g_channels, depth, height, pool=False):
        super().__init__()
        self.at1 = nn.Conv1d(depth, depth, kernel_size=1)
        self.at2 = nn.Conv1d(depth, 1, kernel_size=1)
        self.fc1 = nn.Conv1d(depth, depth, kernel_size=1)
        self.fc2 = nn.Conv1d(depth, channels, kernel_size=1)
        if g_channels is not None:
            self.fc3 = nn.Conv1d(depth, depth, kernel_size=1)
            self.fc4 = nn.Conv1d(depth, g_channels, kernel_size=1)
        self.dropout = nn.Dropout(0.5)
        self.depth, self.height, self.channels, self.g_channels, self.pool = depth, height, channels, g_channels, pool

    def forward_base(self, x, mask, GX, isMixing):
        """x [B,C,Tf,h,w] -> (bias, scale) at the base resolution: [B,ch,Tl,h,w] ([B,ch,Tl,1,1] if pool)."""
        x = X.cl(x)
        B, C, Tf, h, w = x.shape
        if mask.shape[1] != Tf:                                            # :205-207 (two tiny resamplings, as the reference)
            mask = F.adaptive_max_pool1d(mask.unsqueeze(1), Tf).squeeze(1)
            GX = F.adaptive_avg_pool2d(GX.unsqueeze(1), (Tf, None)).squeeze(1)
        if GX.shape[0] != B:                                               # multi-crop testing (:209-211): crops share the features
            n = GX.shape[0] // B
            x = X.cl(x.repeat_interleave(n, dim=0))
            mask = mask.repeat_interleave(n, dim=0)
            B = GX.shape[0]
        mask, GX = mask.contiguous(), GX.contiguous()
        Tl = GX.shape[2]
        P = h * w
        rows = FU.rows_of(x)                                               # [B, Tf*P, C]
        att = FU.linear_rows(FU.linear_rows(rows, self.at1, X.ACT_RELU), self.at2, X.ACT_SIGMOID)      # :216-219
        agg = FU.RewightAggFn.apply(rows.view(B, Tf, P, C), att.view(B, Tf, P), GX, mask)               # :221-225
        if self.pool:                                                      # :227-228
            agg = X.AvgPoolFn.apply(agg.view(B, Tl, h, w, C).permute(0, 4, 1, 2, 3), h, w)
            agg = agg.permute(0, 2, 3, 4, 1).reshape(B, Tl, C)
            h = w = 1
        else:
            agg = agg.view(B, Tl * P, C)
        x1 = FU.linear_rows(agg, self.fc1, X.ACT_RELU)
        if self.pool:
            x1 = self.dropout(x1)
        x1 = FU.from_rows(FU.linear_rows(x1, self.fc2), Tl, h, w)
        if self.g_channels is None:
            return x1, None
        x2 = FU.linear_rows(agg, self.fc3, X.ACT_RELU)
        if self.pool:
            x2 = self.dropout(x2)
        x2 = FU.from_rows(FU.linear_rows(x2, self.fc4, X.ACT_NONE if isMixing else X.ACT_SIGMOID), Tl, h, w)
        return x1, x2

    def forward(self, inp):
        x, lx, mask, gx, i, GX, isMixing = inp
        x1, x2 = self.forward_base(x, mask, GX, isMixing)
        if not self.pool and x.shape[3] != self.height:
            x1 = FU.NearestUpFn.apply(x1, self.height, self.height)
            x2 = FU.NearestUpFn.apply(x2, self.height, self.height) if x2 is not None else None
        return x1 if x2 is None else (x1, x2)


class Gaussian(nn.Module):
    """Temporal alignment weights between fine steps and coarse sample points (x3d_coarse.py:251-286)."""

    def __init__(self, ratio=1):
        super().__init__()
        self.ratio = ratio

    def forward(self, inp):
        meta, mask, gx, tx = inp
        b, b2 = meta.shape[0], gx.shape[0]
        st = meta[:, 0].float()
        if tx is None:                                       # :272-274: no Grid Pool, the coarse steps are a uniform grid 0..Tl-1
            gx = torch.arange(gx.shape[2], device=gx.device, dtype=torch.float32).repeat(b2, 1)
            tx = 1.0
        if b2 != b:                                          # multi-crop testing (:264-266, :279): crop k starts k*step later
            n = b2 // b
            off = meta[:, 3].float().view(-1, 1) * torch.arange(n, device=meta.device, dtype=torch.float32).view(1, -1)
            st = (st.view(-1, 1) + off).reshape(-1)
            mask = mask.repeat_interleave(n, dim=0)
        return FU.GaussianFn.apply(gx.contiguous(), st.contiguous(), mask.contiguous(), float(tx), float(self.ratio))


class MixingLayer(nn.Module):
    """Mixes the four stage-wise shift / scale maps into the one applied at this stage
    (x3d_coarse.py:289-351)."""

    def __init__(self, depth, learned=False, index=0, isLogit=False):
        super().__init__()
        self.learned, self.index, self.isLogit = learned, index, isLogit
        self.in_depth = 432 if isLogit else (24 + 48 + 96 + 192)
        self.range = 1 if isLogit else 4
        self.dropout = nn.Dropout(0.5)
        if learned:
            self.conv_at = nn.Conv1d(self.in_depth, depth, kernel_size=1)
            self.conv_at2 = nn.Conv1d(self.in_depth, depth, kernel_size=1)

    def mix(self, bias, scale, h, w):
        """bias/scale: lists of [B,c_i,Tl,h,w] maps at one common resolution -> (cs, ms) [B,depth,Tl,h,w]."""
        Tl = bias[0].shape[2]
        cs = torch.cat([FU.rows_of(t) for t in bias[:self.range]], dim=2)          # :327
        ms = torch.cat([FU.rows_of(t) for t in scale[:self.range]], dim=2)         # :328
        if not self.learned:
            raise NotImplementedError("learnedMixing=False (one-hot stage select, x3d_coarse.py:339-344) is not built")
        if self.isLogit:
            cs, ms = self.dropout(cs), self.dropout(ms)
        cs = FU.linear_rows(cs, self.conv_at)                                      # :335
        ms = FU.linear_rows(ms, self.conv_at2, X.ACT_SIGMOID)                      # :336
        return FU.from_rows(cs, Tl, h, w), FU.from_rows(ms, Tl, h, w)

    def forward(self, inp):
        x, bias, scale = inp
        h, w = x.shape[3], x.shape[4]
        bias = [FU.resize_map(t, h, w) for t in bias[:self.range]]                 # :312-325
        scale = [FU.resize_map(t, h, w) for t in scale[:self.range]]
        return self.mix(bias, scale, h, w)


class _PoolCfg:
    def __init__(self, training, bn1, bn2):
        self.training, self.bn1, self.bn2 = training, X.BNCfg(bn1), X.BNCfg(bn2)


class GridPoolLayer(nn.Module):
    """Learnable temporal Grid Pool (x3d_coarse.py:355-416): per-interval confidence -> CDF ->
    inverse-transform sample points -> temporal lerp gather.  Returns (x_pooled, cdf)."""

    def __init__(self, ratio, depth):
        super().__init__()
        self.ratio = 4                                       # hard-coded in the reference (:359)
        self.depth = depth
        self.conv1 = nn.Conv3d(depth, depth, kernel_size=3, stride=(self.ratio // 2, 2, 2), padding=1)
        self.bn1 = SubBatchNorm3d(num_splits=1, num_features=depth, affine=True)
        self.conv2 = nn.Conv3d(depth, depth, kernel_size=3, stride=(self.ratio // 2, 2, 2), padding=1)
        self.bn2 = SubBatchNorm3d(num_splits=1, num_features=depth, affine=True)
        self.conv3 = nn.Conv3d(depth, 1, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        self.relu = nn.ReLU(inplace=True)
        self.sigmoid = nn.Sigmoid()

    def confidence(self, x):
        return FU.ConfidenceFn.apply(x, _PoolCfg(self.training, self.bn1, self.bn2), self.conv1.weight, self.conv1.bias,
                                     self.bn1.weight, self.bn1.bias, self.conv2.weight, self.conv2.bias, self.bn2.weight,
                                     self.bn2.bias, self.conv3.weight, self.conv3.bias)

    def forward(self, inp):
        x = X.cl(inp)
        cdf = G.gridpool_cdf(self.confidence(x))             # :379-392
        return G.temporal_sample(x, cdf), cdf                # :394-403


def GridUnpool(inp):
    """Inverse of the Grid Pool re-sampling (x3d_coarse.py:419-451)."""
    x, gx, is_logit = inp
    return G.grid_unpool(x, gx, bool(is_logit), ratio=4)


_REF_CHILD_ORDER = ('pool_1', 'conv1_s', 'conv1_t', 'bn1', 'relu', 'layer1', 'layer2', 'layer3', 'layer4', 'conv5', 'bn5',
                    'rw2', 'rw3', 'rw4', 'rw5', 'rw6', 'mix2', 'mix3', 'mix4', 'mix5', 'gauss', 'avgpool', 'fc1', 'fc2', 'dropout')


# ----------------------------------------------------------------------------------------
class ResNet(_FineResNet):
    """Coarse-stream X3D with Grid Pool after layer1, Multi-stage Fusion of the fine features
    before layer2..layer4 / conv5 and on the logits, Grid Unpool at the end (x3d_coarse.py:455-727)."""

    def __init__(self, block, layers, block_inplanes, n_input_channels=3, feat_depth={}, conv1_t_size=7, conv1_t_stride=1,
                 shortcut_type='B', widen_factor=1.0, dropout=0.5, n_classes=400, base_bn_splits=8, task='class',
                 extract_feat=False, t_pool=None, learnedMixing=False, isMixing=False):
        super().__init__(block, layers, block_inplanes, n_input_channels=n_input_channels, conv1_t_size=conv1_t_size,
                         conv1_t_stride=conv1_t_stride, shortcut_type=shortcut_type, widen_factor=widen_factor,
                         dropout=dropout, n_classes=n_classes, base_bn_splits=base_bn_splits, task=task,
                         extract_feat=extract_feat, global_tower=False, t_downsample=False)
        planes = [(int(a * widen_factor), int(b * widen_factor)) for a, b in block_inplanes]
        self.feat_depth = feat_depth
        self.learnedMixing, self.isMixing, self.t_pool = learnedMixing, isMixing, t_pool
        if t_pool == 'avg':
            self.pool_1 = nn.AvgPool3d((4, 1, 1), stride=(4, 1, 1))
        elif t_pool == 'max':
            self.pool_1 = nn.MaxPool3d((4, 1, 1), stride=(4, 1, 1))
        elif t_pool == 'grid':
            self.pool_1 = GridPoolLayer(ratio=4, depth=planes[0][1])
        self.rw2 = RewightLayer(planes[0][1], planes[0][1], feat_depth['layer1'], height=56)
        self.rw3 = RewightLayer(planes[1][1], planes[1][1], feat_depth['layer2'], height=28)
        self.rw4 = RewightLayer(planes[2][1], planes[2][1], feat_depth['layer3'], height=14)
        self.rw5 = RewightLayer(planes[3][1], planes[3][1], feat_depth['layer4'], height=7)
        self.rw6 = RewightLayer(157, 157, feat_depth['conv5'], height=7, pool=True)
        if isMixing:
            for i in range(4):
                setattr(self, f"mix{i + 2}", MixingLayer(depth=planes[i][1], learned=learnedMixing, index=i))
        self.gauss = Gaussian(ratio=1)
        # registration order of the reference's constructor (x3d_coarse.py:489-555): named_parameters() order is what
        # torch.optim.SGD checkpoints index by, so resuming from / saving to the scripts' optimizer_state_dict needs it
        first = [k for k in _REF_CHILD_ORDER if k in self._modules]
        self._modules = {k: self._modules[k] for k in first + [k for k in self._modules if k not in first]}
        for m in self.modules():                             # same init rule as the reference (:557-561)
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

    def replace_logits(self, n_classes):
        dev = self.fc1.weight.device
        self.fc2 = nn.Linear(2048, n_classes).to(dev)
        self.rw6 = RewightLayer(n_classes, n_classes, self.feat_depth['conv5'], height=7, pool=True).to(dev)

    fusion_streams = True          # run the independent Rewight branches of the fusion block on side streams (CUDA only)

    def _fork_branches(self, rws, feats, feat_masks, GX, flags, first=0):
        """Start rw.forward_base(...) of independent Rewight layers.  Each branch is a chain of small kernels (k=1 Conv1d
        GEMMs over a few thousand rows, the Tf-contraction): launch-latency-bound in sequence, so on CUDA each branch is
        forked onto its own stream; autograd runs a branch's backward on the same stream.  -> handle for _join_branches."""
        if not (self.fusion_streams and feats[0].is_cuda):
            return [(None, rw.forward_base(f, feat_masks, GX, fl)) for rw, f, fl in zip(rws, feats, flags)]
        main = torch.cuda.current_stream()
        streams = X.side_streams(feats[0].device, first + len(rws))[first:]
        pending = []
        for rw, f, fl, s in zip(rws, feats, flags, streams):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                pending.append((s, rw.forward_base(f, feat_masks, GX, fl)))
        return pending

    @staticmethod
    def _join_branches(pending):
        """Wait for forked branches on the current stream and hand their outputs over to it."""
        outs = []
        for s, o in pending:
            if s is not None:
                main = torch.cuda.current_stream()
                main.wait_stream(s)
                for t in o:
                    if t is not None:
                        t.record_stream(main)
            outs.append(o)
        return outs

    def _forward(self, inp):
        x, feat, feat_masks, i, meta = inp
        t_in = x.shape[2]
        x = self.layer1(self._stem(x))                                         # :633-638
        if self.t_pool == 'grid':
            x, gx = self.pool_1(x)                                             # :646-649
            GX = self.gauss([meta, feat_masks, gx, t_in])                      # :650
        else:                                                                  # :640-645, :652 (not used by the shipped scripts)
            gx = None
            if self.t_pool in ('avg', 'max'):
                x = X.cl(self.pool_1(x))                                       # nn.AvgPool3d / nn.MaxPool3d over 4 frames
            elif self.t_pool == 'stride':
                x = X.cl(x[:, :, ::4])
            GX = self.gauss([meta, feat_masks, x, None])
        keys = ('layer1', 'layer2', 'layer3', 'layer4')
        rws = (self.rw2, self.rw3, self.rw4, self.rw5)
        layers = (self.layer2, self.layer3, self.layer4, None)
        # rw6 (the logits' scale / shift, :719-720) depends on the fine features and GX only: it starts here, on its own stream,
        # and is joined at the head (extract_feat returns before the head and never needs it)
        rw6_pending = None if self.extract_feat else self._fork_branches([self.rw6], [feat['conv5']], feat_masks, GX, [False], first=0)
        if self.isMixing:                                                      # :655-679
            maps = self._join_branches(self._fork_branches(rws, [feat[k] for k in keys], feat_masks, GX, [True] * 4, first=1))
            bias, scale = [m[0] for m in maps], [m[1] for m in maps]
            hb, wb = bias[0].shape[3], bias[0].shape[4]
            for j, layer in enumerate(layers):
                c, m = getattr(self, f"mix{j + 2}").mix(bias, scale, hb, wb)
                x = FU.FilmFn.apply(x, m, c)
                if layer is not None:
                    x = layer(x)
        else:                                                                  # :681-699
            for rw, k, layer in zip(rws, keys, layers):
                c, m = rw.forward_base(feat[k], feat_masks, GX, False)
                x = FU.FilmFn.apply(x, m, c)
                if layer is not None:
                    x = layer(x)
        B, _, Tl, H, W = x.shape
        pooled = self._conv5_pool(x, H, W)                                     # :700-702
        if self.task == 'class':
            pooled = pooled.mean(dim=2, keepdim=True)
        if self.extract_feat:
            return pooled
        logits = self._head(pooled)                                            # [B,n_cls,Tl]  :706-716
        (b6, s6), = self._join_branches(rw6_pending)                           # :719-720 (started right after the Gaussian)
        lg = logits.unsqueeze(3).unsqueeze(4)
        x = FU.FilmFn.apply(lg, s6, b6).squeeze(4).squeeze(3)                  # :721
        if self.t_pool != 'grid':
            return x
        x = GridUnpool([x, gx, True])                                          # :724
        return G.linear_upsample_t(x, (x.shape[2] - 1) * 4)                    # :725


def generate_model(x3d_version, **kwargs):
    return ResNet(Bottleneck, get_blocks(x3d_version), get_inplanes(x3d_version), **kwargs)
