"""Autograd functions of the Multi-stage Fusion block and of the Grid Pool confidence branch
over the C ABI (cf_gaussian_*, cf_rewight_agg_*, cf_film_*, cf_nearest_up*, cf_pw_conv with the
tap gather for pool_1.conv1-3).

Everything in the fusion block runs at the 7x7 base resolution of the fine features on
channels-last row tensors [B, R, C]; only FiLM touches the full-resolution coarse activation."""
import torch

from . import x3d_ops as X
from ._lib import call, call_struct, make, ptr, stream_ptr

CL3 = torch.channels_last_3d


def rows_of(x):
    """[B,C,T,H,W] (any layout) -> channels-last rows [B, T*H*W, C] (a view when already channels-last)."""
    x = X.cl(x)
    B, C = x.shape[:2]
    return x.permute(0, 2, 3, 4, 1).reshape(B, -1, C)


def from_rows(rows, T, H, W):
    """rows [B, T*H*W, C] -> logical [B,C,T,H,W] with channels-last strides (a view)."""
    B, _, C = rows.shape
    return rows.view(B, T, H, W, C).permute(0, 4, 1, 2, 3)


# ----------------------------------------------------------------------------------------
class GaussianFn(torch.autograd.Function):
    """Gaussian.forward (x3d_coarse.py:256-286) with tx given: cdf [B,Tl] -> GX [B,Tf,Tl]."""

    @staticmethod
    def forward(ctx, cdf, start, mask, tx, ratio):
        cdf = cdf.contiguous().float()
        start = start.contiguous().float()
        mask = mask.contiguous().float()
        B, Tl = cdf.shape
        Tf = mask.shape[1]
        gx = torch.empty(B, Tf, Tl, device=cdf.device, dtype=torch.float32)
        call("cf_gaussian_fwd", ptr(cdf), ptr(start), ptr(mask), ptr(gx), B, Tf, Tl, float(tx), float(ratio), stream_ptr())
        ctx.save_for_backward(cdf, start, mask)
        ctx.misc = (B, Tf, Tl, float(tx), float(ratio))
        return gx

    @staticmethod
    def backward(ctx, dgx):
        cdf, start, mask = ctx.saved_tensors
        B, Tf, Tl, tx, ratio = ctx.misc
        dcdf = torch.zeros_like(cdf)
        call("cf_gaussian_bwd", ptr(cdf), ptr(start), ptr(mask), ptr(dgx.contiguous()), ptr(dcdf), B, Tf, Tl, tx, ratio,
             stream_ptr())
        return dcdf, None, None, None, None


class RewightAggFn(torch.autograd.Function):
    """Attention-filtered, Gaussian-aligned aggregation over fine time (x3d_coarse.py:221-225):
    x [B,Tf,P,C], att [B,Tf,P], gx [B,Tf,Tl], mask [B,Tf] -> agg [B,Tl,P,C]."""

    @staticmethod
    def forward(ctx, x, att, gx, mask):
        x, att, gx, mask = x.contiguous(), att.contiguous(), gx.contiguous(), mask.contiguous().float()
        B, Tf, P, C = x.shape
        Tl = gx.shape[2]
        agg = torch.empty(B, Tl, P, C, device=x.device, dtype=torch.float32)
        den = torch.empty(B, Tl, P, device=x.device, dtype=torch.float32)
        call_struct("cf_rewight_agg_fwd", make("cf_rewight_args", x=x, att=att, gx=gx, mask=mask, agg=agg, den=den, B=B, C=C,
                                               Tf=Tf, Tl=Tl, P=P))
        ctx.save_for_backward(x, att, gx, mask, agg, den)
        return agg

    @staticmethod
    def backward(ctx, dagg):
        x, att, gx, mask, agg, den = ctx.saved_tensors
        B, Tf, P, C = x.shape
        Tl = gx.shape[2]
        dagg = dagg.contiguous()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        datt = torch.empty_like(att)
        dgx = torch.zeros_like(gx)
        call_struct("cf_rewight_agg_bwd", make("cf_rewight_bwd_args", x=x, att=att, gx=gx, mask=mask, agg=agg, den=den,
                                               dagg=dagg, dx=dx, datt=datt, dgx=dgx, B=B, C=C, Tf=Tf, Tl=Tl, P=P))
        return dx, datt, dgx, None


class FilmFn(torch.autograd.Function):
    """x * scale + shift with scale/shift at a base resolution dividing (H,W)
    (x3d_coarse.py:664,669,674,679,721).  x [B,C,T,H,W]; scale, shift [B,C,T,Hb,Wb]."""

    @staticmethod
    def forward(ctx, x, scale, shift):
        x, scale, shift = X.cl(x), X.cl(scale), X.cl(shift)
        B, C, T, H, W = x.shape
        Hb, Wb = scale.shape[3], scale.shape[4]
        out = torch.empty_like(x)
        call_struct("cf_film_fwd", make("cf_film_args", x=x, scale=scale, shift=shift, out=out, B=B, C=C, T=T, H=H, W=W,
                                        Hb=Hb, Wb=Wb))
        ctx.save_for_backward(x, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, scale = ctx.saved_tensors
        B, C, T, H, W = x.shape
        Hb, Wb = scale.shape[3], scale.shape[4]
        dout = X.cl(dout)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dscale = torch.empty_like(scale)
        dshift = torch.empty_like(scale)
        call_struct("cf_film_bwd", make("cf_film_bwd_args", dout=dout, x=x, scale=scale, dx=dx, dscale=dscale, dshift=dshift,
                                        B=B, C=C, T=T, H=H, W=W, Hb=Hb, Wb=Wb))
        return dx, dscale, dshift


class NearestUpFn(torch.autograd.Function):
    """Exact nearest replication [B,C,T,Hb,Wb] -> [B,C,T,H,W]: what F.adaptive_max_pool2d computes
    at x3d_coarse.py:214,315,322 when (H,W) are multiples of (Hb,Wb)."""

    @staticmethod
    def forward(ctx, x, H, W):
        x = X.cl(x)
        B, C, T, Hb, Wb = x.shape
        out = X.new_act(B, C, T, H, W, x.device)
        call("cf_nearest_up", ptr(x), ptr(out), B, T, Hb, Wb, H, W, C, stream_ptr())
        ctx.dims = (B, C, T, Hb, Wb, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C, T, Hb, Wb, H, W = ctx.dims
        dout = X.cl(dout)
        dx = X.new_act(B, C, T, Hb, Wb, dout.device)
        call("cf_nearest_up_bwd", ptr(dout), ptr(dx), B, T, Hb, Wb, H, W, C, stream_ptr())
        return dx, None, None


# ----------------------------------------------------------------------------------------
class ConfidenceFn(torch.autograd.Function):
    """Grid Pool confidence branch (x3d_coarse.py:362-366,379-383):
    conv1(3^3,s2,bias)->bn1->relu->conv2(3^3,s2,bias)->bn2->relu->conv3((1,3,3),s(1,2,2),bias)->mean(H,W).
    x [B,C,T,H,W] -> g [B, T/4].  The dense convs run as tap-gathered GEMMs (K = C*27 / C*9)."""

    @staticmethod
    def forward(ctx, x, cfg, w1, c1b, g1, b1, w2, c2b, g2, b2, w3, c3b):
        x = X.cl(x)
        dev = x.device
        B, C, T, H, W = x.shape
        tr = cfg.training
        o = lambda n, k, s, p: (n + 2 * p - k) // s + 1
        T1, H1, W1 = o(T, 3, 2, 1), o(H, 3, 2, 1), o(W, 3, 2, 1)
        T2, H2, W2 = o(T1, 3, 2, 1), o(H1, 3, 2, 1), o(W1, 3, 2, 1)
        H3, W3 = o(H2, 3, 2, 1), o(W2, 3, 2, 1)
        g1g = X.geom(T1, H1, W1, T, H, W, k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1), pos_stride=C, sample_stride=T * H * W * C)
        g2g = X.geom(T2, H2, W2, T1, H1, W1, k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1), pos_stride=C,
                     sample_stride=T1 * H1 * W1 * C)
        g3g = X.geom(T2, H3, W3, T2, H2, W2, k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1), pos_stride=C,
                     sample_stride=T2 * H2 * W2 * C)
        R1, R2 = T1 * H1 * W1, T2 * H2 * W2
        stats = torch.zeros(2, B, C, 2, device=dev, dtype=torch.float64) if tr else None
        smode = X.STATS_SUM_SQ if tr else X.STATS_NONE
        y1 = X.new_act(B, C, T1, H1, W1, dev)
        X.pw_conv(x, w1, y1, B, C * 27, C, g1g, bias=c1b, gather_in=1, stats=stats[0] if tr else None, stats_mode=smode)
        a1, bb1, m1, i1 = X.bn_finalize(stats[0] if tr else None, cfg.bn1, B, C, R1, tr, dev)
        y2 = X.new_act(B, C, T2, H2, W2, dev)
        X.pw_conv(y1, w2, y2, B, C * 27, C, g2g, bias=c2b, gather_in=1, pro=X.PRO_AFFINE_RELU, pro_tabs=(a1, bb1, None),
                  stats=stats[1] if tr else None, stats_mode=smode)
        a2, bb2, m2, i2 = X.bn_finalize(stats[1] if tr else None, cfg.bn2, B, C, R2, tr, dev)
        y3 = X.new_act(B, 1, T2, H3, W3, dev)
        X.pw_conv(y2, w3, y3, B, C * 9, 1, g3g, bias=c3b, gather_in=1, pro=X.PRO_AFFINE_RELU, pro_tabs=(a2, bb2, None))
        g = X.new_act(B, 1, T2, 1, 1, dev)
        call_struct("cf_block_avgpool_fwd", make("cf_pool_args", x=y3, y=g, tab_a=None, tab_b=None, B=B, C=1, T=T2, H=H3, W=W3,
                                                 rh=H3, rw=W3))
        ctx.cfg = cfg
        ctx.dims = (B, C, T, H, W, T1, H1, W1, T2, H2, W2, H3, W3)
        ctx.geoms = (g1g, g2g, g3g)
        ctx.tabs = (a1, bb1, m1, i1, a2, bb2, m2, i2)
        ctx.save_for_backward(x, y1, y2, w1, g1, w2, g2, w3)
        ctx.params = (w1, c1b, g1, b1, w2, c2b, g2, b2, w3, c3b)
        return g.view(B, T2)

    @staticmethod
    def backward(ctx, dg):
        x, y1, y2, w1, g1, w2, g2, w3 = ctx.saved_tensors
        B, C, T, H, W, T1, H1, W1, T2, H2, W2, H3, W3 = ctx.dims
        g1g, g2g, g3g = ctx.geoms
        a1, bb1, m1, i1, a2, bb2, m2, i2 = ctx.tabs
        tr = ctx.cfg.training
        dev = x.device
        R1, R2 = T1 * H1 * W1, T2 * H2 * W2
        (dw1, dc1b, dg1, db1, dw2, dc2b, dg2, db2, dw3, dc3b), rets = X._flat_grads(ctx.params, dev)
        sums = torch.zeros(2, B, C, 2, device=dev, dtype=torch.float64)
        # mean over (H3,W3)
        dy3 = X.new_act(B, 1, T2, H3, W3, dev)
        call_struct("cf_block_avgpool_bwd", make("cf_pool_bwd_args", dy=dg.contiguous().float(), x=None, tab_a=None, tab_b=None,
                                                 dz=dy3, sums=None, B=B, C=1, T=T2, H=H3, W=W3, rh=H3, rw=W3, accumulate=0))
        # conv3
        X.pw_wgrad(dy3, y2, dw3, B, C * 9, 1, g3g, x_mode=X.PRO_AFFINE_RELU, x_tabs=(a2, bb2), dbias=dc3b, gather_in=1)
        dA2 = torch.zeros_like(y2)
        X.pw_conv(dy3, w3, dA2, B, 1, C * 9, g3g, w_sn=1, w_sk=C * 9, scatter_out=1)
        dz2 = torch.empty_like(y2)      # through relu(bn2(.)): mask + (sum dz, sum dz*y2)
        call_struct("cf_block_avgpool_bwd", make("cf_pool_bwd_args", dy=dA2, x=y2, tab_a=a2, tab_b=bb2, dz=dz2, sums=sums[1], B=B,
                                                 C=C, T=T2, H=H2, W=W2, rh=1, rw=1, accumulate=0))
        P2, Q2, R2c = X.bn_bwd_coeffs(sums[1], g2, m2, i2, dg2, db2, B, C, R2, tr)
        # conv2
        X.pw_wgrad(dz2, y1, dw2, B, C * 27, C, g2g, dy2=y2, dy_mode=X.PRO_AFFINE2, dy_tabs=(P2, Q2, R2c),
                   x_mode=X.PRO_AFFINE_RELU, x_tabs=(a1, bb1), dbias=dc2b, gather_in=1)
        dA1 = torch.zeros_like(y1)
        X.pw_conv(dz2, w2, dA1, B, C, C * 27, g2g, w_sn=1, w_sk=C * 27, x2=y2, pro=X.PRO_AFFINE2, pro_tabs=(P2, Q2, R2c),
                  scatter_out=1)
        dz1 = torch.empty_like(y1)
        call_struct("cf_block_avgpool_bwd", make("cf_pool_bwd_args", dy=dA1, x=y1, tab_a=a1, tab_b=bb1, dz=dz1, sums=sums[0], B=B,
                                                 C=C, T=T1, H=H1, W=W1, rh=1, rw=1, accumulate=0))
        P1, Q1, R1c = X.bn_bwd_coeffs(sums[0], g1, m1, i1, dg1, db1, B, C, R1, tr)
        # conv1
        X.pw_wgrad(dz1, x, dw1, B, C * 27, C, g1g, dy2=y1, dy_mode=X.PRO_AFFINE2, dy_tabs=(P1, Q1, R1c), dbias=dc1b,
                   gather_in=1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.zeros_like(x)
            X.pw_conv(dz1, w1, dx, B, C, C * 27, g1g, w_sn=1, w_sk=C * 27, x2=y1, pro=X.PRO_AFFINE2, pro_tabs=(P1, Q1, R1c),
                      scatter_out=1)
        return (dx, None, *rets)


def linear_rows(x, conv, act=X.ACT_NONE):
    """k=1 nn.Conv1d / 1x1x1 conv holder applied to a row tensor."""
    return X.LinearRowsFn.apply(x, conv.weight, conv.bias, act)


class BlockMaxPoolFn(torch.autograd.Function):
    """F.adaptive_max_pool2d as a down-sampler over (H,W) of [B,C,T,H,W] (x3d_coarse.py:315,322)."""

    @staticmethod
    def forward(ctx, x, Ho, Wo):
        x = X.cl(x)
        B, C, T, H, W = x.shape
        out = X.new_act(B, C, T, Ho, Wo, x.device)
        idx = torch.empty(B, T, Ho, Wo, C, device=x.device, dtype=torch.int32)
        call("cf_block_maxpool_fwd", ptr(x), ptr(out), ptr(idx), B, T, H, W, C, H // Ho, W // Wo, stream_ptr())
        ctx.save_for_backward(idx)
        ctx.dims = (B, C, T, H, W, Ho, Wo)
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        B, C, T, H, W, Ho, Wo = ctx.dims
        dout = X.cl(dout)
        dx = X.new_act(B, C, T, H, W, dout.device)
        call("cf_block_maxpool_bwd", ptr(dout), ptr(idx), ptr(dx), B, T, H, W, C, H // Ho, W // Wo, stream_ptr())
        return dx, None, None


def resize_map(x, h, w):
    """What F.adaptive_max_pool2d(x.view(b,c*t,hf,wf),(h,w)) computes at x3d_coarse.py:214,315,322
    when the sizes divide each other: replication up, block max down."""
    hf, wf = x.shape[3], x.shape[4]
    if (hf, wf) == (h, w):
        return x
    if h % hf == 0 and w % wf == 0:
        return NearestUpFn.apply(x, h, w)
    if hf % h == 0 and wf % w == 0:
        return BlockMaxPoolFn.apply(x, h, w)
    raise NotImplementedError(f"adaptive_max_pool2d {hf}x{wf} -> {h}x{w}: sizes must divide each other")
