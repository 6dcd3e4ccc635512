"""CPU oracle for the Coarse-Fine X3D hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (fp32, CPU-runnable) *restatement* of the reference's
algorithm, written functionally over a state-dict with the reference's key layout.
It is the checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package (``coarse_fine_networks_b200``) never imports it and has no CPU fallback.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by importing the
reference itself (``/root/reference/x3d_fine.py``, ``x3d_coarse.py``, ``interp1d.py``)
in the build container with ``tests/golden/make_golden.py``; ``tests/test_oracle.py``
checks every function below against those vectors (and, when ``/root/reference`` is
present, against the live reference modules).

Every function cites the reference lines it follows (paths relative to /root/reference).
Where the reference composes generic ATen ops (meshgrid + 5-D ``grid_sample``, 6-D
broadcast products, ``adaptive_max_pool2d`` used as a nearest up-sampler) the oracle uses
the closed forms verified in SURVEY.md section 8(a); these are the forms the CUDA kernels
implement.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

BN_EPS = 1e-5          # nn.BatchNorm3d default, x3d_fine.py:27-29
BN_MOMENTUM = 0.1
FP32_EPS = 1.1920928955078125e-07   # kept in every dtype so that an fp64 run referees the fp32 semantics


# ----------------------------------------------------------------------------------------
# X3D building blocks
# ----------------------------------------------------------------------------------------

def sub_batchnorm(x: Tensor, sd: StateDict, p: str, train: bool, num_splits: int = 1,
                  new_stats: Optional[dict] = None) -> Tensor:
    """SubBatchNorm3d.forward, x3d_fine.py:51-62 (same x3d_coarse.py:49-60).

    train: batch statistics over groups of n/num_splits samples (view trick at :54-56),
    no affine inside BN, then ``*weight`` and ``+bias`` (:59-61).  eval: running stats of
    ``<p>.bn`` (:58).  ``new_stats`` (optional dict) receives the momentum-updated running
    statistics of ``split_bn`` the reference would hold after this call.
    """
    n, c, t, h, w = x.shape
    if train:
        xs = x.reshape(n // num_splits, c * num_splits, t, h, w)
        mean = xs.mean(dim=(0, 2, 3, 4))
        var = xs.var(dim=(0, 2, 3, 4), unbiased=False)
        y = (xs - mean.view(1, -1, 1, 1, 1)) * torch.rsqrt(var.view(1, -1, 1, 1, 1) + BN_EPS)
        y = y.reshape(n, c, t, h, w)
        if new_stats is not None:
            cnt = xs.numel() // xs.shape[1]
            rm = sd[p + ".split_bn.running_mean"]
            rv = sd[p + ".split_bn.running_var"]
            new_stats[p + ".split_bn.running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
            new_stats[p + ".split_bn.running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * (cnt / max(cnt - 1, 1))
    else:
        rm = sd[p + ".bn.running_mean"].view(1, -1, 1, 1, 1)
        rv = sd[p + ".bn.running_var"].view(1, -1, 1, 1, 1)
        y = (x - rm) * torch.rsqrt(rv + BN_EPS)
    y = y * sd[p + ".weight"].view(1, -1, 1, 1, 1)
    y = y + sd[p + ".bias"].view(1, -1, 1, 1, 1)
    return y


def aggregate_bn_stats(sd: StateDict, p: str, num_splits: int) -> Tuple[Tensor, Tensor]:
    """SubBatchNorm3d._get_aggregated_mean_std / aggregate_stats, x3d_fine.py:31-47."""
    means = sd[p + ".split_bn.running_mean"].view(num_splits, -1)
    stds = sd[p + ".split_bn.running_var"].view(num_splits, -1)
    mean = means.sum(0) / num_splits
    std = stds.sum(0) / num_splits + ((means - mean) ** 2).sum(0) / num_splits
    return mean, std


def swish(x: Tensor) -> Tensor:
    """SwishEfficient, x3d_fine.py:74-86: x*sigmoid(x); autograd gives the same backward."""
    return x * torch.sigmoid(x)


def bottleneck(x: Tensor, sd: StateDict, p: str, stride: int, index: int, train: bool,
               num_splits: int = 1, new_stats: Optional[dict] = None) -> Tensor:
    """Bottleneck.forward, x3d_fine.py:146-175 / x3d_coarse.py:143-172."""
    out = F.conv3d(x, sd[p + ".conv1.weight"])
    out = F.relu(sub_batchnorm(out, sd, p + ".bn1", train, num_splits, new_stats))
    ce = out.shape[1]
    out = F.conv3d(out, sd[p + ".conv2.weight"], stride=(1, stride, stride), padding=1, groups=ce)
    out = sub_batchnorm(out, sd, p + ".bn2", train, num_splits, new_stats)
    if index % 2 == 0:                                   # SE, x3d_fine.py:157-163
        se = out.mean(dim=(2, 3, 4), keepdim=True)
        se = F.relu(F.conv3d(se, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]))
        se = torch.sigmoid(F.conv3d(se, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"]))
        out = out * se
    out = swish(out)
    out = F.conv3d(out, sd[p + ".conv3.weight"])
    out = sub_batchnorm(out, sd, p + ".bn3", train, num_splits, new_stats)
    if (p + ".downsample.0.weight") in sd:               # _make_layer, x3d_fine.py:277-288
        res = F.conv3d(x, sd[p + ".downsample.0.weight"], stride=(1, stride, stride))
        res = sub_batchnorm(res, sd, p + ".downsample.1", train, num_splits, new_stats)
    else:
        res = x
    return F.relu(out + res)


def _n_blocks(sd: StateDict, layer: str) -> int:
    n = 0
    while f"{layer}.{n}.conv1.weight" in sd:
        n += 1
    return n


def stage(x: Tensor, sd: StateDict, layer: str, train: bool, num_splits: int = 1,
          new_stats: Optional[dict] = None) -> Tensor:
    """One ``layerK`` Sequential: first block stride 2 + downsample, rest stride 1
    (x3d_fine.py:277-306)."""
    for i in range(_n_blocks(sd, layer)):
        x = bottleneck(x, sd, f"{layer}.{i}", 2 if i == 0 else 1, i, train, num_splits, new_stats)
    return x


def stem(x: Tensor, sd: StateDict, train: bool, num_splits: int = 1,
         new_stats: Optional[dict] = None) -> Tensor:
    """conv1_s -> conv1_t -> bn1 -> relu, x3d_fine.py:334-337."""
    x = F.conv3d(x, sd["conv1_s.weight"], stride=(1, 2, 2), padding=(0, 1, 1))
    x = F.conv3d(x, sd["conv1_t.weight"], padding=(2, 0, 0), groups=x.shape[1])
    return F.relu(sub_batchnorm(x, sd, "bn1", train, num_splits, new_stats))


def head(x: Tensor, sd: StateDict) -> Tensor:
    """avgpool(None,1,1) -> fc1 -> relu -> (dropout p=0) -> fc2, 'loc' task,
    x3d_fine.py:366-380.  x: [B,432,T,h,w] -> [B,n_classes,T].  Dropout is the identity
    here (parity protocol, SURVEY 8(a) finding 1)."""
    x = x.mean(dim=(3, 4))                                  # [B,C,T]
    x = F.relu(torch.einsum("oc,bct->bot", sd["fc1.weight"].flatten(1), x))
    return torch.einsum("oc,bct->bot", sd["fc2.weight"], x) + sd["fc2.bias"].view(1, -1, 1)


def fine_forward(sd: StateDict, x: Tensor, train: bool, num_splits: int = 1,
                 global_tower: bool = False, new_stats: Optional[dict] = None):
    """x3d_fine.ResNet.forward, x3d_fine.py:331-382 ('loc' task, dropout=0)."""
    feats = {}
    x = stem(x, sd, train, num_splits, new_stats)
    for name in ("layer1", "layer2", "layer3", "layer4"):
        x = stage(x, sd, name, train, num_splits, new_stats)
        if global_tower:
            feats[name] = F.adaptive_avg_pool3d(x, (None, 7, 7))     # :345-354
    x = F.relu(sub_batchnorm(F.conv3d(x, sd["conv5.weight"]), sd, "bn5", train, num_splits, new_stats))
    if global_tower:
        feats["conv5"] = F.adaptive_avg_pool3d(x, (None, 7, 7))      # :360
        return feats
    return head(x, sd)


# ----------------------------------------------------------------------------------------
# Grid Pool / Grid Unpool / Interp1d
# ----------------------------------------------------------------------------------------

def gridpool_cdf(g: Tensor) -> Tensor:
    """Pre-sigmoid per-interval confidences g [B,n] -> CDF [B,n+1] in [0,1].
    x3d_coarse.py:384-392: sigma(0.5 g); p = 1-sigma; p/(sum p + 1e-16); cumsum; prepend 0."""
    s = torch.sigmoid(g * 5e-1)
    p = 1.0 - s
    p = p / (p.sum(dim=1, keepdim=True) + 1e-16)
    c = torch.cumsum(p, dim=1)
    return torch.cat([torch.zeros_like(c[:, :1]), c], dim=1)


def sample_coords(cdf: Tensor, t_in: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Normalised grid coordinate and un-normalisation exactly as the reference + ATen do
    in fp32: g = (cdf-0.5)*2 (x3d_coarse.py:394, :440); z = ((g+1)/2)*(T-1)
    (grid_sampler_unnormalize, align_corners=True, called from x3d_coarse.py:403/445).
    Returns (z, i0 = floor(z) as int64  -- the bit-exact "frame-index bins" --, w1 = z-i0)."""
    g = (cdf - 0.5) * 2
    z = ((g + 1) / 2) * (t_in - 1)
    i0f = torch.floor(z)
    return z, i0f.to(torch.int64), z - i0f


def temporal_lerp(x: Tensor, cdf: Tensor) -> Tensor:
    """Closed form of meshgrid+stack+F.grid_sample(align_corners=True, zero padding) when
    the spatial grid sits on pixel centres (x3d_coarse.py:396-403, :442-445):
    out[b,c,k] = (1-w1) x[b,c,i0] + w1 x[b,c,i0+1]; a corner outside [0,T-1] contributes 0.
    x: [B,C,T,*spatial]; cdf: [B,K] -> [B,C,K,*spatial].  Differentiable w.r.t. x and cdf."""
    b, c, t = x.shape[:3]
    _, i0, _ = sample_coords(cdf.detach(), t)
    g = (cdf - 0.5) * 2
    z = ((g + 1) / 2) * (t - 1)
    w1 = z - i0.to(z.dtype)
    w0 = 1.0 - w1
    i1 = i0 + 1
    v0 = ((i0 >= 0) & (i0 <= t - 1)).to(x.dtype)
    v1 = ((i1 >= 0) & (i1 <= t - 1)).to(x.dtype)
    sp = x.shape[3:]
    k = cdf.shape[1]
    idx_shape = (b, 1, k) + (1,) * len(sp)
    exp_shape = (b, c, k) + tuple(sp)
    x0 = torch.gather(x, 2, i0.clamp(0, t - 1).view(idx_shape).expand(exp_shape))
    x1 = torch.gather(x, 2, i1.clamp(0, t - 1).view(idx_shape).expand(exp_shape))
    return (w0 * v0).view(idx_shape) * x0 + (w1 * v1).view(idx_shape) * x1


def gridpool_confidence(sd: StateDict, p: str, x: Tensor, train: bool,
                        new_stats: Optional[dict] = None) -> Tensor:
    """Confidence branch of GridPoolLayer, x3d_coarse.py:362-366, 379-383:
    conv(3^3,s2,bias)->bn->relu ->conv(3^3,s2,bias)->bn->relu ->conv((1,3,3),s(1,2,2),bias)
    -> mean over (H,W).  Returns g [B, T/4]."""
    g = F.conv3d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=2, padding=1)
    g = F.relu(sub_batchnorm(g, sd, p + ".bn1", train, 1, new_stats))
    g = F.conv3d(g, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], stride=2, padding=1)
    g = F.relu(sub_batchnorm(g, sd, p + ".bn2", train, 1, new_stats))
    g = F.conv3d(g, sd[p + ".conv3.weight"], sd[p + ".conv3.bias"], stride=(1, 2, 2), padding=(0, 1, 1))
    return g.mean(dim=(3, 4)).squeeze(1)


def gridpool_layer(sd: StateDict, p: str, x: Tensor, train: bool,
                   new_stats: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """GridPoolLayer.forward, x3d_coarse.py:373-416 -> (x_pooled [B,C,T/4+1,H,W], cdf [B,T/4+1])."""
    cdf = gridpool_cdf(gridpool_confidence(sd, p, x, train, new_stats))
    return temporal_lerp(x, cdf), cdf


def interp1d(x: Tensor, y: Tensor, xnew: Tensor) -> Tuple[Tensor, Tensor]:
    """Interp1d.forward for 2-D inputs, interp1d.py:100-141.  Returns (ynew, ind).
    ind = clamp(searchsorted(x, xnew) - 1, 0, N-2) (:100-110, right=False);
    slope = (y[1:]-y[:-1]) / (eps + x[1:]-x[:-1]) with eps = fp32 machine eps (:37,133-137);
    ynew = y[ind] + slope[ind]*(xnew - x[ind]) (:140-141)."""
    eps = FP32_EPS            # torch.finfo(torch.float32).eps: the reference runs in fp32 (interp1d.py:37)
    ind = torch.searchsorted(x.detach().contiguous(), xnew.detach().contiguous()) - 1
    ind = ind.clamp(0, x.shape[1] - 2)
    slopes = (y[:, 1:] - y[:, :-1]) / (eps + (x[:, 1:] - x[:, :-1]))
    ynew = torch.gather(y, 1, ind) + torch.gather(slopes, 1, ind) * (xnew - torch.gather(x, 1, ind))
    return ynew, ind


def inverse_cdf(cdf: Tensor) -> Tuple[Tensor, Tensor]:
    """x3d_coarse.py:435-438: mid = arange(N)/(N-1); gx_ = Interp1d()(cdf, mid, mid)."""
    n = cdf.shape[1]
    mid = torch.arange(n, dtype=cdf.dtype, device=cdf.device)
    mid = (mid / (n - 1.0)).view(1, -1).repeat(cdf.shape[0], 1)
    return interp1d(cdf, mid, mid)


def linear_upsample_t(x: Tensor, t_out: int) -> Tensor:
    """Linear interpolation along dim 2 with align_corners=True (F.interpolate 'linear' at
    x3d_coarse.py:725; the temporal part of 'trilinear' at :449 whose spatial size is
    unchanged, hence identity in (H,W))."""
    t_in = x.shape[2]
    scale = (t_in - 1) / (t_out - 1) if t_out > 1 else 0.0
    u = torch.arange(t_out, dtype=x.dtype, device=x.device)
    src = u * torch.tensor(scale, dtype=x.dtype)
    j0 = src.to(torch.int64).clamp(max=t_in - 1)
    j1 = (j0 + 1).clamp(max=t_in - 1)
    lam = (src - j0.to(x.dtype))
    shp = (1, 1, t_out) + (1,) * (x.dim() - 3)
    return x.index_select(2, j0) * (1 - lam).view(shp) + x.index_select(2, j1) * lam.view(shp)


def gridunpool(x: Tensor, cdf: Tensor, is_logit: bool) -> Tensor:
    """GridUnpool, x3d_coarse.py:419-451.  logit mode: x [B,C,T'] -> [B,C,T'];
    feature mode: x [B,C,T',H,W] -> [B,C,4T',H,W] (adds the trilinear interpolate of :449)."""
    inv, _ = inverse_cdf(cdf)
    y = temporal_lerp(x, inv)
    if not is_logit:
        y = linear_upsample_t(y, x.shape[2] * 4)
    return y


# ----------------------------------------------------------------------------------------
# Multi-stage fusion
# ----------------------------------------------------------------------------------------

def gaussian(meta: Tensor, mask: Tensor, cdf: Tensor, tx: int, ratio: float = 1.0) -> Tensor:
    """Gaussian.forward with tx given (grid pooling), x3d_coarse.py:256-286.
    mu[b,k] = (cdf[b,k]*tx + meta[b,0]) / ratio (:270,275); sigma_b = sum_t mask / 8 (:278);
    f = exp(-(t-mu)^2 / (2 sigma^2 + 1e-16)) (:280-282), divided by (max_t f + 1e-16) (:283).
    -> [B, Tf, K]."""
    b, tf = mask.shape
    st = meta[:, 0].to(cdf.dtype)
    mu = (cdf * tx + st.view(b, 1)) / ratio                      # [B,K]
    std = mask.sum(dim=1) / 8.0                                  # [B]
    t = torch.arange(tf, dtype=cdf.dtype, device=cdf.device).view(1, tf, 1)
    d = t - mu.view(b, 1, -1)
    f = torch.exp(-(d ** 2) / (2 * (std ** 2).view(b, 1, 1) + 1e-16))
    return f / (f.max(dim=1, keepdim=True)[0] + 1e-16)


def _conv1d_k1(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """nn.Conv1d(kernel_size=1) on [B,C,L]."""
    return torch.einsum("oc,bcl->bol", sd[p + ".weight"].squeeze(-1), x) + sd[p + ".bias"].view(1, -1, 1)


def nearest_up(x: Tensor, h: int) -> Tensor:
    """F.adaptive_max_pool2d used as an up-sampler (x3d_coarse.py:214,315,322): output bin
    i covers input rows floor(i*hin/h) .. ceil((i+1)*hin/h)-1, a single row when h is a
    multiple of hin -> exact nearest replication."""
    hin = x.shape[-1]
    if hin == h:
        return x
    assert h % hin == 0
    r = h // hin
    return x.repeat_interleave(r, dim=-2).repeat_interleave(r, dim=-1)


def rewight(sd: StateDict, p: str, x: Tensor, mask: Tensor, GX: Tensor, height: int,
            pool: bool, is_mixing: bool) -> Tuple[Tensor, Tensor]:
    """RewightLayer.forward, x3d_coarse.py:199-247, evaluated at the 7x7 resolution of the
    fine features and replicated to ``height`` afterwards (exact: every op after the
    nearest up-sampling at :214 is per-pixel).  x: [B,C,Tf,7,7]; mask [B,Tf]; GX [B,Tf,Tl].
    Returns (bias, scale): [B,ch,Tl,h,h] (h=1 when ``pool``).  Dropout = identity."""
    b, c, tf, h, w = x.shape
    tl = GX.shape[2]
    at = F.relu(_conv1d_k1(x.reshape(b, c, -1), sd, p + ".at1"))
    at = torch.sigmoid(_conv1d_k1(at, sd, p + ".at2")).view(b, tf, h, w)            # :216-219
    A = at.view(b, tf, 1, h, w) * GX.view(b, tf, tl, 1, 1)                            # :221
    m = mask.view(b, tf, 1, 1, 1)
    den = (A * m).sum(dim=1, keepdim=True) + 1e-6                                     # :224
    wgt = A * m / den                                                                 # [B,Tf,Tl,h,w]
    agg = torch.einsum("bcthw,btkhw->bckhw", x, wgt)                                  # :222,225
    if pool:
        agg = agg.mean(dim=(3, 4), keepdim=True)                                      # :227-228
    bb, cc, kk, hh, ww = agg.shape
    flat = agg.reshape(bb, cc, -1)
    x1 = _conv1d_k1(F.relu(_conv1d_k1(flat, sd, p + ".fc1")), sd, p + ".fc2").view(bb, -1, kk, hh, ww)
    x2 = _conv1d_k1(F.relu(_conv1d_k1(flat, sd, p + ".fc3")), sd, p + ".fc4").view(bb, -1, kk, hh, ww)
    if not is_mixing:
        x2 = torch.sigmoid(x2)                                                        # :244-245
    if not pool:
        x1, x2 = nearest_up(x1, height), nearest_up(x2, height)
    return x1, x2


def mixing(sd: StateDict, p: str, biases: Sequence[Tensor], scales: Sequence[Tensor], h: int) -> Tuple[Tensor, Tensor]:
    """MixingLayer.forward (learned=True, isLogit=False), x3d_coarse.py:307-351: resize the
    four bias / scale maps to (h,h) (:312-325), concat to 360 channels (:327-328), k=1 Conv1d
    (:335) and sigmoid(k=1 Conv1d) (:336).  Inputs here are the *base-resolution* maps (7x7 in
    the network: every map the reference feeds in is an exact replication of its 7x7 base, so
    max-pooling it down or replicating it up commutes with the per-pixel convs); the result is
    replicated to h."""
    cs = torch.cat(list(biases), dim=1)
    ms = torch.cat(list(scales), dim=1)
    b, _, tl, hh, ww = cs.shape
    c = _conv1d_k1(cs.reshape(b, cs.shape[1], -1), sd, p + ".conv_at").view(b, -1, tl, hh, ww)
    m = torch.sigmoid(_conv1d_k1(ms.reshape(b, ms.shape[1], -1), sd, p + ".conv_at2")).view(b, -1, tl, hh, ww)
    return nearest_up(c, h), nearest_up(m, h)


def coarse_forward(sd: StateDict, x: Tensor, feat: Dict[str, Tensor], feat_masks: Tensor,
                   meta: Tensor, train: bool, num_splits: int = 1,
                   new_stats: Optional[dict] = None, return_aux: bool = False):
    """x3d_coarse.ResNet.forward, t_pool='grid', isMixing=True, learnedMixing=True, 'loc'
    task, dropout=0 (x3d_coarse.py:628-727)."""
    t_in = x.shape[2]
    x = stem(x, sd, train, num_splits, new_stats)                                     # :633-636
    x = stage(x, sd, "layer1", train, num_splits, new_stats)                          # :638
    x, cdf = gridpool_layer(sd, "pool_1", x, train, new_stats)                        # :646-649
    GX = gaussian(meta, feat_masks, cdf, t_in)                                        # :650
    rw = {}
    for name, key in (("rw2", "layer1"), ("rw3", "layer2"), ("rw4", "layer3"), ("rw5", "layer4")):
        rw[name] = rewight(sd, name, feat[key], feat_masks, GX, 7, False, True)       # :656-659 (kept at 7x7)
    biases = [rw[k][0] for k in ("rw2", "rw3", "rw4", "rw5")]
    scales = [rw[k][1] for k in ("rw2", "rw3", "rw4", "rw5")]
    for mix, layer in (("mix2", "layer2"), ("mix3", "layer3"), ("mix4", "layer4"), ("mix5", None)):
        c, m = mixing(sd, mix, biases, scales, x.shape[-1])                           # :663,668,673,678
        x = x * m + c
        if layer is not None:
            x = stage(x, sd, layer, train, num_splits, new_stats)
    x = F.relu(sub_batchnorm(F.conv3d(x, sd["conv5.weight"]), sd, "bn5", train, num_splits, new_stats))
    x = head(x, sd)                                                                   # [B,n_cls,Tl]  :702-716
    b6, s6 = rewight(sd, "rw6", feat["conv5"], feat_masks, GX, 1, True, False)        # :719-720
    x = x * s6.flatten(2) + b6.flatten(2)                                             # :721
    x = gridunpool(x, cdf, True)                                                      # :724
    out = linear_upsample_t(x, (x.shape[2] - 1) * 4)                                  # :725
    if return_aux:
        return out, {"cdf": cdf, "GX": GX}
    return out


# ----------------------------------------------------------------------------------------
# Evaluation
# ----------------------------------------------------------------------------------------

def average_precision(scores: Tensor, targets: Tensor, weights: Optional[Tensor] = None) -> Tensor:
    """APMeter.value(), apmeter.py:98-136: per class, sort the scores descending (:117), gather the targets (:118),
    tp = cumsum(truth [* weight]) (:125-128), rg = 1..N or cumsum(weight) (:109,122), precision = tp / rg (:131),
    ap = sum(precision[truth]) / max(sum(truth), 1) (:134).  scores [N,K] float, targets [N,K] 0/1, weights [N] or None."""
    n, k = scores.shape
    ap = torch.zeros(k)
    rg0 = torch.arange(1, n + 1, dtype=torch.float32)
    for c in range(k):
        _, ind = torch.sort(scores[:, c], 0, True)
        truth = targets[:, c][ind].float()
        if weights is not None:
            w = weights[ind].float()
            tp = (truth * w).cumsum(0)
            rg = w.cumsum(0)
        else:
            tp = truth.cumsum(0)
            rg = rg0
        precision = tp / rg
        ap[c] = precision[truth.bool()].sum() / max(float(truth.sum()), 1.0)
    return ap


def localize_samples(probs: Tensor, labels: Tensor, valid_t: int) -> Tuple[Tensor, Tensor]:
    """The 25-point sampling of the Charades localisation protocol, train_coarse_fineFEAT.py:249-253:
    p1 = probs[:, :valid_t]; sc = valid_t / 25.; p1[:, 1::int(sc)][:, :25] (same for the labels).  probs / labels [C,TL]."""
    sc = valid_t / 25.0
    step = int(sc)
    return probs[:, :valid_t][:, 1::step][:, :25], labels[:, :valid_t][:, 1::step][:, :25]
