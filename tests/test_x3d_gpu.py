"""GPU parity of the X3D conv-stack kernels (through the C ABI) against plain PyTorch fp32 on
CPU (kernel level), the oracle and the golden vectors of the reference (module / net level)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cf_oracle as O
from synth import fill_state_dict, synth_state_dict, synth_tensor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CL3 = torch.channels_last_3d


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= atol + rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.fixture(scope="module")
def X():
    from coarse_fine_networks_b200 import x3d_ops
    return x3d_ops


def rows(x):
    """[B,C,T,H,W] -> channels-last cuda tensor"""
    return x.cuda().contiguous(memory_format=CL3)


# ------------------------------------------------------------------------------ kernel level
@pytest.mark.parametrize("K,N", [(24, 54), (54, 24), (108, 48), (432, 192), (192, 432), (7, 5), (432, 2048)])
def test_pw_conv_plain_and_stats(X, K, N):
    B, T, H, W = 2, 3, 9, 7                                 # R = 189: ragged last row tile
    x = synth_tensor((B, K, T, H, W), 1)
    w = synth_tensor((N, K), 2, 0.1)
    ref = torch.einsum("nk,bkthw->bnthw", w, x)
    xc = rows(x)
    y = X.new_act(B, N, T, H, W, "cuda")
    stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.float64)
    X.pw_conv(xc, w.cuda(), y, B, K, N, X.geom(T, H, W), stats=stats, stats_mode=X.STATS_SUM_SQ)
    close(y, ref, rtol=1e-5, atol=1e-5, what="y")
    close(stats[..., 0], ref.double().sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4, what="sum")
    close(stats[..., 1], (ref.double() ** 2).sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4, what="sumsq")
    # transposed weights = data gradient
    g = synth_tensor((B, N, T, H, W), 3)
    dx = X.new_act(B, K, T, H, W, "cuda")
    X.pw_conv(rows(g), w.cuda(), dx, B, N, K, X.geom(T, H, W), w_sn=1, w_sk=K)
    close(dx, torch.einsum("nk,bnthw->bkthw", w, g), rtol=2e-5, atol=1e-5, what="dgrad")
    # weight gradient
    dw = torch.zeros(N, K, device="cuda")
    db = torch.zeros(N, device="cuda")
    X.pw_wgrad(rows(g), xc, dw, B, K, N, X.geom(T, H, W), dbias=db)
    close(dw, torch.einsum("bnthw,bkthw->nk", g, x), rtol=1e-5, atol=1e-4, what="wgrad")
    close(db, g.sum(dim=(0, 2, 3, 4)), rtol=1e-5, atol=1e-4, what="dbias")


def test_pw_conv_prologues_epilogues(X):
    B, K, N, T, H, W = 2, 54, 24, 2, 6, 5
    x, x2 = synth_tensor((B, K, T, H, W), 11), synth_tensor((B, K, T, H, W), 12)
    w = synth_tensor((N, K), 13, 0.2)
    ta, tb, tc = synth_tensor((B, K), 14), synth_tensor((B, K), 15), synth_tensor((B, K), 16)
    aux = synth_tensor((B, N, T, H, W), 17)
    ea, eb = synth_tensor((B, N), 18), synth_tensor((B, N), 19)
    bias = synth_tensor((N,), 20)
    v = lambda t: t.view(B, -1, 1, 1, 1)
    conv = lambda z: torch.einsum("nk,bkthw->bnthw", w, z)
    g = X.geom(T, H, W)
    cu = lambda t: t.cuda()
    sw = lambda z: z * torch.sigmoid(z)
    cases = {
        X.PRO_AFFINE: v(ta) * x + v(tb),
        X.PRO_AFFINE_RELU: F.relu(v(ta) * x + v(tb)),
        X.PRO_AFFINE_SWISH: sw(v(ta) * x + v(tb)),
        X.PRO_AFFINE2: v(ta) * x + v(tb) * x2 + v(tc),
    }
    for mode, xin in cases.items():
        y = X.new_act(B, N, T, H, W, "cuda")
        X.pw_conv(rows(x), cu(w), y, B, K, N, g, x2=rows(x2), pro=mode, pro_tabs=(cu(ta), cu(tb), cu(tc)))
        close(y, conv(xin), rtol=1e-5, atol=1e-5, what=f"pro {mode}")
    base = conv(x)
    pre = v(ea) * aux + v(eb)
    sg = torch.sigmoid(pre)
    epis = {
        X.EPI_RELU: F.relu(base + bias.view(1, -1, 1, 1, 1)),
        X.EPI_DRELU: base * (pre > 0),
        X.EPI_DSWISH: base * (sg * (1 + pre * (1 - sg))),
        X.EPI_ADD_AUX: base + aux,
    }
    for mode, ref in epis.items():
        y = X.new_act(B, N, T, H, W, "cuda")
        stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.float64)
        X.pw_conv(rows(x), cu(w), y, B, K, N, g, bias=cu(bias) if mode == X.EPI_RELU else None, epi=mode, aux=rows(aux),
                  epi_tabs=(cu(ea), cu(eb)), stats=stats, stats_mode=X.STATS_SUM_AUX)
        close(y, ref, rtol=1e-5, atol=1e-5, what=f"epi {mode}")
        close(stats[..., 1], (ref.double() * aux.double()).sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4, what=f"epi {mode} sum*aux")
    # wgrad with both prologues
    dy, dy2 = synth_tensor((B, N, T, H, W), 21), synth_tensor((B, N, T, H, W), 22)
    da, db, dc = synth_tensor((B, N), 23), synth_tensor((B, N), 24), synth_tensor((B, N), 25)
    vn = lambda t: t.view(B, -1, 1, 1, 1)
    dyy = vn(da) * dy + vn(db) * dy2 + vn(dc)
    dw = torch.zeros(N, K, device="cuda")
    X.pw_wgrad(rows(dy), rows(x), dw, B, K, N, g, dy2=rows(dy2), dy_mode=X.PRO_AFFINE2, dy_tabs=(cu(da), cu(db), cu(dc)),
               x_mode=X.PRO_AFFINE_SWISH, x_tabs=(cu(ta), cu(tb)))
    close(dw, torch.einsum("bnthw,bkthw->nk", dyy, sw(v(ta) * x + v(tb))), rtol=1e-5, atol=1e-4, what="wgrad pro")


@pytest.mark.parametrize("stride,dims", [(1, (2, 24, 48, 2, 9, 8)), (2, (2, 24, 48, 2, 9, 8)),
                                         # enough rows for the tensor-core weight gradient with gathered rows; 54 channels = 8-byte path
                                         (2, (3, 54, 24, 8, 30, 28)), (2, (2, 24, 108, 9, 27, 29))])
def test_pw_conv_strided_gather_scatter(X, stride, dims):
    B, K, N, T, H, W = dims
    x = synth_tensor((B, K, T, H, W), 31)
    w = synth_tensor((N, K, 1, 1, 1), 32, 0.2)
    ref = F.conv3d(x, w, stride=(1, stride, stride))
    Ho, Wo = ref.shape[3], ref.shape[4]
    g = X.geom(T, Ho, Wo, T, H, W, s=(1, stride, stride), pos_stride=K, ch_stride=1, sample_stride=T * H * W * K)
    y = X.new_act(B, N, T, Ho, Wo, "cuda")
    X.pw_conv(rows(x), w.cuda(), y, B, K, N, g, gather_in=1)
    close(y, ref, rtol=1e-5, atol=1e-5, what="fwd")
    gy = synth_tensor(tuple(ref.shape), 33)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    F.conv3d(xr, wr, stride=(1, stride, stride)).backward(gy)
    dx = torch.full((B, K, T, H, W), 0.5, device="cuda").contiguous(memory_format=CL3)
    X.pw_conv(rows(gy), w.cuda(), dx, B, N, K, g, w_sn=1, w_sk=K, scatter_out=1, accumulate=1)
    close(dx, xr.grad + 0.5 * (xr.grad != 0) + 0.5 * (xr.grad == 0), rtol=1e-5, atol=1e-5, what="scatter dgrad")
    dw = torch.zeros(N, K, device="cuda")
    X.pw_wgrad(rows(gy), rows(x), dw, B, K, N, g, gather_in=1)
    close(dw, wr.grad.flatten(1), rtol=1e-5, atol=1e-4, what="wgrad gather")


def test_dense_conv_through_tap_gather(X):
    """conv1_s (NCTHW input, 1x3x3 s2) and a pool_1-style 3x3x3 stride-2 conv with bias."""
    B, T, H, W = 2, 3, 12, 10
    x = synth_tensor((B, 3, T, H, W), 41)
    w = synth_tensor((24, 3, 1, 3, 3), 42, 0.3)
    ref = F.conv3d(x, w, stride=(1, 2, 2), padding=(0, 1, 1))
    Ho, Wo = ref.shape[3], ref.shape[4]
    g = X.geom(T, Ho, Wo, T, H, W, k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1), pos_stride=1, ch_stride=T * H * W,
               sample_stride=3 * T * H * W)
    y = X.new_act(B, 24, T, Ho, Wo, "cuda")
    X.pw_conv(x.cuda(), w.cuda(), y, B, 27, 24, g, gather_in=1)
    close(y, ref, rtol=1e-5, atol=1e-5, what="conv1_s")
    gy = synth_tensor(tuple(ref.shape), 43)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    F.conv3d(xr, wr, stride=(1, 2, 2), padding=(0, 1, 1)).backward(gy)
    dw = torch.zeros(24, 27, device="cuda")
    X.pw_wgrad(rows(gy), x.cuda(), dw, B, 27, 24, g, gather_in=1)
    close(dw, wr.grad.flatten(1), rtol=1e-5, atol=1e-4, what="conv1_s wgrad")
    dx = torch.zeros(B, 3, T, H, W, device="cuda")
    X.pw_conv(rows(gy), w.cuda(), dx, B, 24, 27, g, w_sn=1, w_sk=27, scatter_out=1)
    close(dx, xr.grad, rtol=1e-5, atol=1e-5, what="conv1_s dgrad")
    # the specialised stem kernels (x3d_stem.cu) on a temporal WINDOW of a longer clip (the coarse stream reads frames
    # a..b of the fine stream's clip in place), odd sizes, more than one 64-position tile
    Tf, a0, Tw, H2, W2 = 7, 2, 4, 23, 17
    xf = synth_tensor((B, 3, Tf, H2, W2), 44)
    xw = xf[:, :, a0:a0 + Tw]
    ref2 = F.conv3d(xw, w, stride=(1, 2, 2), padding=(0, 1, 1))
    Ho2, Wo2 = ref2.shape[3], ref2.shape[4]
    xc = xf.cuda()
    xv = xc[:, :, a0:a0 + Tw]
    g2 = X.geom(Tw, Ho2, Wo2, Tw, H2, W2, k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1), pos_stride=1, ch_stride=xv.stride(1),
                sample_stride=xv.stride(0))
    y2 = X.new_act(B, 24, Tw, Ho2, Wo2, "cuda")
    X.pw_conv(xv, w.cuda(), y2, B, 27, 24, g2, gather_in=1)
    close(y2, ref2, rtol=1e-5, atol=1e-5, what="conv1_s window")
    gy2 = synth_tensor(tuple(ref2.shape), 45)
    xr2, wr2 = xw.clone().requires_grad_(True), w.clone().requires_grad_(True)
    F.conv3d(xr2, wr2, stride=(1, 2, 2), padding=(0, 1, 1)).backward(gy2)
    dw2 = torch.zeros(24, 27, device="cuda")
    X.pw_wgrad(rows(gy2), xv, dw2, B, 27, 24, g2, gather_in=1)
    close(dw2, wr2.grad.flatten(1), rtol=1e-5, atol=1e-4, what="conv1_s window wgrad")
    # channels-last 3x3x3 stride (2,2,2) conv, 8 -> 8 channels
    C = 8
    x = synth_tensor((B, C, 6, 9, 9), 44)
    w = synth_tensor((C, C, 3, 3, 3), 45, 0.2)
    bias = synth_tensor((C,), 46)
    ref = F.conv3d(x, w, bias, stride=2, padding=1)
    To, Ho, Wo = ref.shape[2:]
    g = X.geom(To, Ho, Wo, 6, 9, 9, k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1), pos_stride=C, ch_stride=1, sample_stride=6 * 81 * C)
    y = X.new_act(B, C, To, Ho, Wo, "cuda")
    X.pw_conv(rows(x), w.cuda(), y, B, C * 27, C, g, gather_in=1, bias=bias.cuda())
    close(y, ref, rtol=1e-5, atol=1e-5, what="3x3x3 s2")
    gy = synth_tensor(tuple(ref.shape), 47)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    F.conv3d(xr, wr, bias, stride=2, padding=1).backward(gy)
    dx = torch.zeros(B, C, 6, 9, 9, device="cuda").contiguous(memory_format=CL3)
    X.pw_conv(rows(gy), w.cuda(), dx, B, C, C * 27, g, w_sn=1, w_sk=C * 27, scatter_out=1)
    close(dx, xr.grad, rtol=1e-5, atol=1e-5, what="3x3x3 dgrad")
    dw = torch.zeros(C, C * 27, device="cuda")
    db = torch.zeros(C, device="cuda")
    X.pw_wgrad(rows(gy), rows(x), dw, B, C * 27, C, g, gather_in=1, dbias=db)
    close(dw, wr.grad.flatten(1), rtol=1e-5, atol=1e-4, what="3x3x3 wgrad")
    close(db, gy.sum(dim=(0, 2, 3, 4)), rtol=1e-5, atol=1e-4, what="3x3x3 dbias")


@pytest.mark.parametrize("C,stride,T,H,W", [(54, 1, 4, 9, 10), (54, 2, 4, 9, 10), (108, 2, 4, 9, 10), (24, 1, 4, 9, 10), (7, 2, 4, 9, 10),
                                            # plane-marching kernels (x3d_dw3.cu): tiles of 8 x 14 / 8 x 7, several T segments, ragged H
                                            (54, 1, 9, 11, 28), (108, 1, 7, 14, 14), (216, 1, 13, 7, 7), (54, 1, 1, 56, 56),
                                            (54, 1, 5, 8, 42),
                                            # stride-2 plane-marching kernels (x3d_dw3s2.cu): output tiles 4 x 14 / 4 x 7
                                            (54, 2, 6, 18, 28), (108, 2, 9, 14, 14), (54, 2, 3, 112, 56), (54, 2, 5, 13, 27)])
def test_depthwise_fwd_bwd(X, C, stride, T, H, W):
    B = 2
    x = synth_tensor((B, C, T, H, W), 51)
    w = synth_tensor((C, 1, 3, 3, 3), 52, 0.3)
    ta, tb = synth_tensor((B, C), 53), synth_tensor((B, C), 54)
    v = lambda t: t.view(B, -1, 1, 1, 1)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    act = F.relu(v(ta) * xr + v(tb))
    ref = F.conv3d(act, wr, stride=(1, stride, stride), padding=1, groups=C)
    Ho, Wo = ref.shape[3], ref.shape[4]
    g = X.geom(T, Ho, Wo, T, H, W, k=(3, 3, 3), s=(1, stride, stride), p=(1, 1, 1))
    y = X.new_act(B, C, T, Ho, Wo, "cuda")
    stats = torch.zeros(B, C, 2, device="cuda", dtype=torch.float64)
    X.dw_call("cf_dw_conv_fwd", rows(x), w.cuda(), y, B, C, g, pro=X.PRO_AFFINE_RELU, pro_tabs=(ta.cuda(), tb.cuda(), None),
              stats=stats, stats_mode=X.STATS_SUM_SQ)
    close(y, ref, rtol=1e-5, atol=1e-5, what="dw fwd")
    close(stats[..., 0], ref.double().sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4, what="dw sum")
    close(stats[..., 1], (ref.double() ** 2).sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4, what="dw sumsq")
    # backward: dy = P*d + Q*y2 + R, then through conv and the relu
    d, y2 = synth_tensor(tuple(ref.shape), 55), synth_tensor(tuple(ref.shape), 56)
    P, Q, Rr = synth_tensor((B, C), 57), synth_tensor((B, C), 58), synth_tensor((B, C), 59)
    dy = v(P) * d + v(Q) * y2 + v(Rr)
    ref.backward(dy)
    pre = v(ta) * x + v(tb)
    dz_ref = xr.grad / v(ta)                                 # gradient w.r.t. the BN output (before *ta)
    dz = X.new_act(B, C, T, H, W, "cuda")
    sums = torch.zeros(B, C, 2, device="cuda", dtype=torch.float64)
    X.dw_call("cf_dw_conv_dgrad", rows(d), w.cuda(), dz, B, C, g, x2=rows(y2), pro=X.PRO_AFFINE2,
              pro_tabs=(P.cuda(), Q.cuda(), Rr.cuda()), aux=rows(x), epi=X.EPI_DRELU, epi_tabs=(ta.cuda(), tb.cuda()),
              stats=sums, stats_mode=X.STATS_SUM_AUX)
    close(dz, dz_ref, rtol=1e-4, atol=1e-4, what="dw dgrad")
    close(sums[..., 0], dz_ref.double().sum(dim=(2, 3, 4)), rtol=1e-4, atol=1e-3, what="dw dgrad sum")
    close(sums[..., 1], (dz_ref.double() * x.double()).sum(dim=(2, 3, 4)), rtol=1e-4, atol=1e-3, what="dw dgrad sum*aux")
    dw = torch.zeros(C, 27, device="cuda")
    X.dw_call("cf_dw_conv_wgrad", rows(d), w.cuda(), dw, B, C, g, x2=rows(y2), pro=X.PRO_AFFINE2,
              pro_tabs=(P.cuda(), Q.cuda(), Rr.cuda()), aux=rows(x), epi_tabs=(ta.cuda(), tb.cuda()))
    close(dw, wr.grad.flatten(1), rtol=1e-4, atol=1e-4, what="dw wgrad")
    # data gradient + weight gradient from ONE call (dw_out): one pass over (d, y2, x) on the stride-1 plane-marching kernel,
    # the two kernels back to back elsewhere; dw_out is accumulated into (+=)
    dz2 = X.new_act(B, C, T, H, W, "cuda")
    sums2 = torch.zeros(B, C, 2, device="cuda", dtype=torch.float64)
    dwf = torch.ones(C, 27, device="cuda")
    X.dw_call("cf_dw_conv_dgrad", rows(d), w.cuda(), dz2, B, C, g, x2=rows(y2), pro=X.PRO_AFFINE2,
              pro_tabs=(P.cuda(), Q.cuda(), Rr.cuda()), aux=rows(x), epi=X.EPI_DRELU, epi_tabs=(ta.cuda(), tb.cuda()),
              stats=sums2, stats_mode=X.STATS_SUM_AUX, dw_out=dwf)
    close(dz2, dz_ref, rtol=1e-4, atol=1e-4, what="fused dw dgrad")
    close(sums2[..., 1], (dz_ref.double() * x.double()).sum(dim=(2, 3, 4)), rtol=1e-4, atol=1e-3, what="fused dw dgrad sum*aux")
    close(dwf - 1.0, wr.grad.flatten(1), rtol=1e-4, atol=2e-4, what="fused dw wgrad")


def test_depthwise_temporal_5x1x1(X):
    B, C, T, H, W = 2, 24, 7, 5, 6
    x = synth_tensor((B, C, T, H, W), 61)
    w = synth_tensor((C, 1, 5, 1, 1), 62, 0.4)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv3d(xr, wr, padding=(2, 0, 0), groups=C)
    g = X.geom(T, H, W, k=(5, 1, 1), p=(2, 0, 0))
    y = X.new_act(B, C, T, H, W, "cuda")
    X.dw_call("cf_dw_conv_fwd", rows(x), w.cuda(), y, B, C, g)
    close(y, ref, rtol=1e-5, atol=1e-5, what="conv1_t")
    gy = synth_tensor(tuple(ref.shape), 63)
    ref.backward(gy)
    dx = X.new_act(B, C, T, H, W, "cuda")
    X.dw_call("cf_dw_conv_dgrad", rows(gy), w.cuda(), dx, B, C, g)
    close(dx, xr.grad, rtol=1e-5, atol=1e-5, what="conv1_t dgrad")
    dw = torch.zeros(C, 5, device="cuda")
    X.dw_call("cf_dw_conv_wgrad", rows(gy), w.cuda(), dw, B, C, g, aux=rows(x))
    close(dw, wr.grad.flatten(1), rtol=1e-5, atol=1e-4, what="conv1_t wgrad")
    # data gradient + weight gradient from one march (dw_out), with the BatchNorm-backward prologue the stem uses,
    # T long enough for several T segments
    T2 = 40
    x = synth_tensor((B, C, T2, H, W), 64)
    d, y2 = synth_tensor((B, C, T2, H, W), 65), synth_tensor((B, C, T2, H, W), 66)
    P, Q, Rr = synth_tensor((B, C), 67), synth_tensor((B, C), 68), synth_tensor((B, C), 69)
    v = lambda t: t.view(B, -1, 1, 1, 1)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    F.conv3d(xr, wr, padding=(2, 0, 0), groups=C).backward(v(P) * d + v(Q) * y2 + v(Rr))
    g2 = X.geom(T2, H, W, k=(5, 1, 1), p=(2, 0, 0))
    dx2 = X.new_act(B, C, T2, H, W, "cuda")
    dwf = torch.ones(C, 5, device="cuda")
    X.dw_call("cf_dw_conv_dgrad", rows(d), w.cuda(), dx2, B, C, g2, x2=rows(y2), pro=X.PRO_AFFINE2,
              pro_tabs=(P.cuda(), Q.cuda(), Rr.cuda()), aux=rows(x), dw_out=dwf)
    close(dx2, xr.grad, rtol=1e-4, atol=1e-4, what="conv1_t fused dgrad")
    close(dwf - 1.0, wr.grad.flatten(1), rtol=1e-4, atol=2e-3, what="conv1_t fused wgrad")


# ------------------------------------------------------------------------------ module level
def _build_bottleneck(g, splits):
    from coarse_fine_networks_b200 import x3d_fine as M
    stride, index = int(g["stride"]), int(g["index"])
    inp, planes = 8, (18, 8)
    down = None
    if stride != 1:
        down = torch.nn.Sequential(M.conv1x1x1(inp, planes[1], stride), M.SubBatchNorm3d(num_splits=splits, num_features=planes[1], affine=True))
    return M.Bottleneck(inp, planes, stride=stride, downsample=down, index=index, base_bn_splits=splits)


@pytest.mark.parametrize("name", ["bottleneck_s2_se", "bottleneck_s1", "bottleneck_s1_se_split2"])
def test_bottleneck_golden(name):
    g = load(name)
    splits = int(g["splits"])
    m = _build_bottleneck(g, splits)
    fill_state_dict(m, seed=62)                        # the state the reference started from
    m.cuda().train()
    x = g["x"].cuda().requires_grad_(True)
    out = m(x)
    close(out, g["out"], rtol=2e-5, atol=2e-5, what="out")
    out.backward(g["gout"].cuda())
    close(x.grad, g["dx"], rtol=5e-4, atol=2e-5, what="dx")
    named = dict(m.named_parameters())
    for k, gr in sub(g, "grad/").items():
        close(named[k].grad, gr, rtol=5e-4, atol=2e-5, what=f"grad {k}")
    after = sub(g, "sd_after/")
    sd = m.state_dict()
    for k in after:
        if "split_bn.running" in k:
            close(sd[k], after[k], rtol=1e-5, atol=1e-6, what=k)
        if k.endswith("split_bn.num_batches_tracked"):
            assert int(sd[k]) == int(after[k])
    m.aggregate = [mod.aggregate_stats() for mod in m.modules() if hasattr(mod, "aggregate_stats")]
    m.eval()
    with torch.no_grad():
        out_eval = m(g["x"].cuda())
    close(out_eval, g["out_eval"], rtol=2e-5, atol=2e-5, what="eval out")


def test_bottleneck_real_widths_vs_oracle():
    from coarse_fine_networks_b200 import x3d_fine as M
    for (cin, planes, stride, index) in [(24, (54, 24), 1, 1), (24, (108, 48), 2, 0), (96, (216, 96), 1, 2)]:
        down = None
        if stride != 1 or cin != planes[1]:
            down = torch.nn.Sequential(M.conv1x1x1(cin, planes[1], stride), M.SubBatchNorm3d(num_splits=1, num_features=planes[1], affine=True))
        m = M.Bottleneck(cin, planes, stride=stride, downsample=down, index=index, base_bn_splits=1)
        fill_state_dict(m, seed=7)
        sd = {"b." + k: v.clone() for k, v in m.state_dict().items()}
        params = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
        x = synth_tensor((2, cin, 4, 14, 14), 8)
        xr = x.clone().requires_grad_(True)
        ref = O.bottleneck(xr, sd, "b", stride, index, True, 1)
        gout = synth_tensor(tuple(ref.shape), 9)
        ref.backward(gout)
        m.cuda().train()
        xc = x.cuda().requires_grad_(True)
        out = m(xc)
        close(out, ref, rtol=5e-5, atol=5e-5, what="out")
        out.backward(gout.cuda())
        close(xc.grad, xr.grad, rtol=1e-3, atol=5e-5, what="dx")
        for k, p in m.named_parameters():
            close(p.grad, params["b." + k].grad, rtol=1e-3, atol=5e-5, what=f"grad {k}")


@pytest.mark.parametrize("stride,grad_out", [(1, True), (2, True), (1, False)])
def test_bottleneck_fused_stage_pool_vs_oracle(stride, grad_out):
    """Stage-final block of the global tower: the residual-join kernel also emits adaptive_avg_pool3d(out, (None,7,7))
    (x3d_fine.py:345-354); forward and backward (gradient arriving through both outputs, or through the features only)
    against the oracle's block followed by torch's pooling."""
    import torch.nn.functional as F
    from coarse_fine_networks_b200 import x3d_fine as M
    cin, planes, index = 24, (54, 24), 0
    down = None
    if stride != 1:
        down = torch.nn.Sequential(M.conv1x1x1(cin, planes[1], stride), M.SubBatchNorm3d(num_splits=1, num_features=planes[1], affine=True))
    m = M.Bottleneck(cin, planes, stride=stride, downsample=down, index=index, base_bn_splits=1)
    fill_state_dict(m, seed=17)
    sd = {"b." + k: v.clone() for k, v in m.state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    x = synth_tensor((2, cin, 3, 28 * stride, 28 * stride), 18)
    xr = x.clone().requires_grad_(True)
    ref = O.bottleneck(xr, sd, "b", stride, index, True, 1)
    ref_p = F.adaptive_avg_pool3d(ref, (None, 7, 7))
    gout, gp = synth_tensor(tuple(ref.shape), 19), synth_tensor(tuple(ref_p.shape), 20)
    ((ref * gout).sum() * (1.0 if grad_out else 0.0) + (ref_p * gp).sum()).backward()
    m.cuda().train()
    xc = x.cuda().requires_grad_(True)
    out, pooled = m(xc, pool=(4, 4))
    close(out, ref, rtol=5e-5, atol=5e-5, what="out")
    close(pooled, ref_p, rtol=5e-5, atol=5e-5, what="pooled")
    loss = (pooled * gp.cuda()).sum()
    if grad_out:
        loss = loss + (out * gout.cuda()).sum()
    loss.backward()
    close(xc.grad, xr.grad, rtol=1e-3, atol=5e-5, what="dx")
    for k, p in m.named_parameters():
        close(p.grad, params["b." + k].grad, rtol=1e-3, atol=5e-5, what=f"grad {k}")


def test_standalone_bn_and_swish():
    from coarse_fine_networks_b200 import x3d_fine as M
    bn = M.SubBatchNorm3d(num_splits=2, num_features=6, affine=True)
    fill_state_dict(bn, seed=3)
    sd = {"bn." + k: v.clone() for k, v in bn.state_dict().items()}
    w, b = sd["bn.weight"].requires_grad_(True), sd["bn.bias"].requires_grad_(True)
    x = synth_tensor((4, 6, 3, 5, 5), 4)
    xr = x.clone().requires_grad_(True)
    ref = O.sub_batchnorm(xr, sd, "bn", True, 2)
    gout = synth_tensor(tuple(ref.shape), 5)
    ref.backward(gout)
    bn.cuda().train()
    xc = x.cuda().requires_grad_(True)
    out = bn(xc)
    close(out, ref, rtol=1e-5, atol=1e-5, what="bn out")
    out.backward(gout.cuda())
    close(xc.grad, xr.grad, rtol=1e-4, atol=1e-5, what="bn dx")
    close(bn.weight.grad, w.grad, rtol=1e-4, atol=1e-4, what="bn dgamma")
    close(bn.bias.grad, b.grad, rtol=1e-4, atol=1e-4, what="bn dbeta")
    xs = x.cuda().requires_grad_(True)
    ys = M.Swish()(xs)
    ys.backward(gout.cuda())
    xr2 = x.clone().requires_grad_(True)
    O.swish(xr2).backward(gout)
    close(ys, O.swish(x), rtol=1e-6, atol=1e-6, what="swish")
    close(xs.grad, xr2.grad, rtol=1e-5, atol=1e-6, what="swish bwd")


# ------------------------------------------------------------------------------ whole fine net
def test_fine_net_golden():
    from coarse_fine_networks_b200 import x3d_fine as M
    g = load("fine_net")
    m = M.generate_model("S", n_classes=10, task="loc", base_bn_splits=1, dropout=0.0)
    fill_state_dict(m, seed=72)
    m.cuda()
    x = synth_tensor((2, 3, 4, 64, 64), seed=73)
    m.eval()
    with torch.no_grad():
        out_eval = m([x.cuda(), None])
    close(out_eval, g["out_eval"], rtol=2e-4, atol=2e-5, what="eval logits")
    m.train()
    xg = x.cuda().requires_grad_(True)
    out = m([xg, None])
    close(out, g["out_train"], rtol=1e-3, atol=1e-4, what="train logits")      # north-star tolerance: 1e-3 rel
    out.backward(synth_tensor(tuple(out.shape), seed=74).cuda())
    # sum over all positions of a signed gradient: heavy cancellation on top of the whole-net conditioning (SURVEY 8(a)
    # finding 3); run to run this lands between 0.5 % and 3 % of the scale (order of the statistics atomics)
    close(xg.grad.sum(dim=(2, 3, 4)), g["dx_sum"], rtol=5e-2, atol=1e-3, what="dx")
    named = dict(m.named_parameters())
    for k, gr in sub(g, "grad/").items():
        close(named[k].grad, gr, rtol=2e-2, atol=1e-4, what=f"grad {k}")      # whole-net grads: SURVEY 8(a) finding 3


def test_fine_net_global_tower_vs_oracle():
    from coarse_fine_networks_b200 import x3d_fine as M
    m = M.generate_model("M", n_classes=157, task="loc", base_bn_splits=1, dropout=0.0, global_tower=True)
    fill_state_dict(m, seed=5)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = synth_tensor((1, 3, 4, 224, 224), seed=6)
    with torch.no_grad():
        ref = O.fine_forward(sd, x, False, global_tower=True)
    m.cuda().eval()
    with torch.no_grad():
        feats, masks = m([x.cuda(), None])
    assert masks is None
    for k in ("layer1", "layer2", "layer3", "layer4", "conv5"):
        assert feats[k].shape == ref[k].shape
        close(feats[k], ref[k], rtol=2e-4, atol=2e-5, what=k)


def test_x3d_xl_variant_runs_and_matches_the_oracle():
    """generate_model('XL') (x3d_fine.py:388-402: blocks [5,10,25,15], widths 72/162/306/630 -> 32/72/136/280): none of the
    widths is a multiple of 54, so every depthwise conv takes the general kernels and the GEMMs other tilings; train-mode
    logits against the oracle on a small clip (train mode: the key-hashed synthetic running statistics do not describe a
    normalised net), and the backward runs."""
    from coarse_fine_networks_b200 import x3d_fine as M
    m = M.generate_model("XL", n_classes=7, task="loc", base_bn_splits=1, dropout=0.0)
    sd = synth_state_dict(m.state_dict(), 91)
    m.load_state_dict(sd, strict=True)
    x = synth_tensor((2, 3, 4, 64, 64), 92)
    with torch.no_grad():
        ref = O.fine_forward(sd, x, True)
    m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    o2 = m([xg, None])
    assert o2.shape == (2, 7, 4)
    err = (o2.detach().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 1e-3, err
    o2.square().mean().backward()
    assert torch.isfinite(xg.grad).all() and all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
