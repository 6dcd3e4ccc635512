"""Autograd wrappers of the Grid Pool / Grid Unpool kernels (C ABI: cf_gridpool_*,
cf_sample_bins, cf_temporal_gather_*, cf_inverse_cdf_*, cf_linear_bins)."""
import torch

from ._lib import call, lib, ptr, stream_ptr


def _layout(x):
    """[B,C,T,*sp] tensor -> (dense tensor, outer, outer_per_b, T, inner, channels_last?)."""
    B, C, T = x.shape[:3]
    inner_sp = 1
    for s in x.shape[3:]:
        inner_sp *= s
    if x.dim() == 5 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last_3d):
        return x, B, 1, T, inner_sp * C, True
    return x.contiguous(), B * C, C, T, inner_sp, False


def _alloc_like(x, t_new, cl):
    shape = list(x.shape)
    shape[2] = t_new
    if cl:
        return torch.empty(shape, dtype=x.dtype, device=x.device, memory_format=torch.channels_last_3d)
    return torch.empty(shape, dtype=x.dtype, device=x.device)


def _match(g, cl):
    return g.contiguous(memory_format=torch.channels_last_3d) if cl else g.contiguous()


class GridPoolCdf(torch.autograd.Function):
    """g [B,n] -> cdf [B,n+1] (x3d_coarse.py:384-392)."""

    @staticmethod
    def forward(ctx, g):
        g = g.contiguous().float()
        B, n = g.shape
        cdf = torch.empty(B, n + 1, device=g.device, dtype=torch.float32)
        call("cf_gridpool_cdf_fwd", ptr(g), ptr(cdf), B, n, stream_ptr())
        ctx.save_for_backward(g)
        return cdf

    @staticmethod
    def backward(ctx, dcdf):
        (g,) = ctx.saved_tensors
        B, n = g.shape
        dg = torch.empty_like(g)
        call("cf_gridpool_cdf_bwd", ptr(g), ptr(dcdf.contiguous()), ptr(dg), B, n, stream_ptr())
        return dg


def sample_bins(coord, t_in):
    """coord [..] in [0,1] -> (i0 int32, w1 fp32) frame-index bins for a T=t_in source."""
    coord = coord.detach().contiguous().float()
    i0 = torch.empty(coord.shape, device=coord.device, dtype=torch.int32)
    w1 = torch.empty(coord.shape, device=coord.device, dtype=torch.float32)
    call("cf_sample_bins", ptr(coord), ptr(i0), ptr(w1), coord.numel(), int(t_in), stream_ptr())
    return i0, w1


def linear_bins(t_in, t_out, device):
    i0 = torch.empty(t_out, device=device, dtype=torch.int32)
    w1 = torch.empty(t_out, device=device, dtype=torch.float32)
    call("cf_linear_bins", ptr(i0), ptr(w1), int(t_in), int(t_out), stream_ptr())
    return i0, w1


def _gather_fwd(x, i0, w1, per_batch):
    xd, outer, opb, T, inner, cl = _layout(x)
    K = i0.shape[-1]
    out = _alloc_like(xd, K, cl)
    call("cf_temporal_gather_fwd", ptr(xd), ptr(i0), ptr(w1), ptr(out), outer, opb if per_batch else outer, T, K, inner,
         stream_ptr())
    return xd, out, (outer, opb if per_batch else outer, T, K, inner, cl)


def _gather_bwd_x(gout, i0, w1, meta, x_like):
    outer, opb, T, K, inner, cl = meta
    gout = _match(gout, cl)
    dx = torch.empty_like(x_like)
    nb = (outer + opb - 1) // opb
    ws_bytes = int(lib.cf_temporal_gather_bwd_ws_bytes(nb, T, K))
    ws = torch.empty(ws_bytes, device=gout.device, dtype=torch.uint8)
    call("cf_temporal_gather_bwd_x", ptr(gout), ptr(i0), ptr(w1), ptr(dx), ptr(ws), ws_bytes, outer, opb, T, K, inner,
         stream_ptr())
    return gout, dx


class TemporalSample(torch.autograd.Function):
    """F.grid_sample along T at per-sample coordinates in [0,1] (x3d_coarse.py:394-403,
    440-445): x [B,C,T,*sp], coord [B,K] -> [B,C,K,*sp].  Differentiable in x and coord."""

    @staticmethod
    def forward(ctx, x, coord):
        coord = coord.contiguous().float()
        i0, w1 = sample_bins(coord, x.shape[2])
        xd, out, meta = _gather_fwd(x, i0, w1, True)
        ctx.save_for_backward(xd, i0, w1)
        ctx.meta = meta
        return out

    @staticmethod
    def backward(ctx, gout):
        xd, i0, w1 = ctx.saved_tensors
        outer, opb, T, K, inner, cl = ctx.meta
        gout = _match(gout, cl)
        dx = dcoord = None
        if ctx.needs_input_grad[0]:
            _, dx = _gather_bwd_x(gout, i0, w1, ctx.meta, xd)
        if ctx.needs_input_grad[1]:
            dcoord = torch.zeros(i0.shape, device=gout.device, dtype=torch.float64)
            call("cf_temporal_gather_bwd_coord", ptr(gout), ptr(xd), ptr(i0), ptr(dcoord), outer, opb, T, K, inner,
                 float(T - 1), stream_ptr())
            dcoord = dcoord.float()
        return dx, dcoord


class LinearUpsampleT(torch.autograd.Function):
    """F.interpolate(..., mode='linear'|'trilinear', align_corners=True) along T only
    (x3d_coarse.py:449 with unchanged (H,W), and :725)."""

    @staticmethod
    def forward(ctx, x, t_out):
        i0, w1 = linear_bins(x.shape[2], t_out, x.device)
        xd, out, meta = _gather_fwd(x, i0, w1, False)
        ctx.save_for_backward(i0, w1)
        ctx.meta = meta
        ctx.x_shape_like = (xd.shape, xd.stride())
        return out

    @staticmethod
    def backward(ctx, gout):
        i0, w1 = ctx.saved_tensors
        shape, stride = ctx.x_shape_like
        x_like = torch.empty_strided(shape, stride, device=gout.device, dtype=gout.dtype)
        _, dx = _gather_bwd_x(gout, i0, w1, ctx.meta, x_like)
        return dx, None


class InverseCdf(torch.autograd.Function):
    """Interp1d()(cdf, mid, mid) with mid = arange(K)/(K-1)  (x3d_coarse.py:435-438)."""

    @staticmethod
    def forward(ctx, cdf):
        cdf = cdf.contiguous().float()
        B, K = cdf.shape
        inv = torch.empty_like(cdf)
        ind = torch.empty(B, K, device=cdf.device, dtype=torch.int32)
        call("cf_inverse_cdf_fwd", ptr(cdf), ptr(inv), ptr(ind), B, K, stream_ptr())
        ctx.save_for_backward(cdf, ind)
        ctx.mark_non_differentiable(ind)
        return inv, ind

    @staticmethod
    def backward(ctx, dinv, _dind):
        cdf, ind = ctx.saved_tensors
        B, K = cdf.shape
        dcdf = torch.zeros_like(cdf)
        call("cf_inverse_cdf_bwd", ptr(cdf), ptr(ind), ptr(dinv.contiguous()), ptr(dcdf), B, K, stream_ptr())
        return dcdf


def gridpool_cdf(g):
    return GridPoolCdf.apply(g)


def temporal_sample(x, coord):
    return TemporalSample.apply(x, coord)


def linear_upsample_t(x, t_out):
    return LinearUpsampleT.apply(x, int(t_out))


def inverse_cdf(cdf):
    return InverseCdf.apply(cdf)


def grid_unpool(x, cdf, is_logit, ratio=4):
    """GridUnpool (x3d_coarse.py:419-451)."""
    inv, _ = inverse_cdf(cdf)
    y = temporal_sample(x, inv)
    if not is_logit:
        y = linear_upsample_t(y, x.shape[2] * ratio)
    return y
