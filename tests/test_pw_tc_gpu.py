"""tcgen05 (3xTF32) pointwise GEMM against an fp64 reference and against the fp32 CUDA-core kernel."""
import pytest
import torch
import torch.nn.functional as F

from synth import synth_tensor

pytestmark = pytest.mark.gpu
CL3 = torch.channels_last_3d


@pytest.fixture(scope="module")
def X():
    import __graft_entry__ as ge
    ge.build()
    from coarse_fine_networks_b200 import x3d_ops
    return x3d_ops


def rows(x):
    return x.cuda().contiguous(memory_format=CL3)


def relerr(a, ref):
    return ((a.detach().cpu().double() - ref).abs().max() / ref.abs().max()).item()


@pytest.mark.parametrize("K,N", [(24, 54), (54, 24), (48, 108), (108, 48), (96, 216), (216, 96), (192, 432), (432, 192),
                                 (432, 2048), (2048, 157), (360, 24), (7, 5), (33, 130), (8, 16)])
def test_tc_forward_dgrad_stats_vs_fp64(X, K, N):
    B, T, H, W = 2, 2, 13, 11                               # R = 286 rows per sample: 3 row tiles, ragged tail
    x = synth_tensor((B, K, T, H, W), 1)
    w = synth_tensor((N, K), 2, 0.1)
    ref = torch.einsum("nk,bkthw->bnthw", w.double(), x.double())
    g = X.geom(T, H, W)
    out = {}
    for tc in (True, False):
        y = X.new_act(B, N, T, H, W, "cuda")
        stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.float64)
        X.pw_conv(rows(x), w.cuda(), y, B, K, N, g, stats=stats, stats_mode=X.STATS_SUM_SQ, tc=tc)
        out[tc] = (relerr(y, ref), relerr(stats[..., 0], ref.sum(dim=(2, 3, 4))), relerr(stats[..., 1], (ref ** 2).sum(dim=(2, 3, 4))))
    # 3xTF32 must be as accurate as an fp32 FMA chain (well inside the 1e-3 parity bar; single-pass TF32 would give ~1e-3)
    # measured: 4e-7 (K=24) .. 4e-6 (K=432) .. 1.6e-5 (K=2048; grows with the number of sequential fp32 accumulations in
    # TMEM, ~770 MMAs deep there); the fp32 FMA chain gives 1e-7 .. 1.5e-6.  All far inside the 1e-3 parity bar.
    tol = 1e-5 if K <= 512 else 3e-5
    assert out[True][0] <= tol, out
    assert out[True][1] <= 2 * tol and out[True][2] <= 3 * tol, out
    gy = synth_tensor((B, N, T, H, W), 3)
    dref = torch.einsum("nk,bnthw->bkthw", w.double(), gy.double())
    dx = X.new_act(B, K, T, H, W, "cuda")
    X.pw_conv(rows(gy), w.cuda(), dx, B, N, K, g, w_sn=1, w_sk=K, tc=True)
    assert relerr(dx, dref) <= (1e-5 if N <= 512 else 3e-5)


def test_tc_prologues_epilogues_accumulate(X):
    B, K, N, T, H, W = 2, 54, 24, 2, 12, 11
    x, x2 = synth_tensor((B, K, T, H, W), 11), synth_tensor((B, K, T, H, W), 12)
    w = synth_tensor((N, K), 13, 0.2)
    ta, tb, tcc = synth_tensor((B, K), 14), synth_tensor((B, K), 15), synth_tensor((B, K), 16)
    aux = synth_tensor((B, N, T, H, W), 17)
    ea, eb = synth_tensor((B, N), 18), synth_tensor((B, N), 19)
    bias = synth_tensor((N,), 20)
    v = lambda t: t.double().view(B, -1, 1, 1, 1)
    conv = lambda z: torch.einsum("nk,bkthw->bnthw", w.double(), z.double())
    g = X.geom(T, H, W)
    cu = lambda t: t.cuda()
    sw = lambda z: z * torch.sigmoid(z)
    xd, x2d, auxd = x.double(), x2.double(), aux.double()
    cases = {X.PRO_AFFINE: v(ta) * xd + v(tb), X.PRO_AFFINE_RELU: F.relu(v(ta) * xd + v(tb)),
             X.PRO_AFFINE_SWISH: sw(v(ta) * xd + v(tb)), X.PRO_AFFINE2: v(ta) * xd + v(tb) * x2d + v(tcc)}
    for mode, xin in cases.items():
        y = X.new_act(B, N, T, H, W, "cuda")
        X.pw_conv(rows(x), cu(w), y, B, K, N, g, x2=rows(x2), pro=mode, pro_tabs=(cu(ta), cu(tb), cu(tcc)), tc=True)
        assert relerr(y, conv(xin)) <= 5e-6, f"pro {mode}"
    base = conv(x)
    pre = v(ea) * auxd + v(eb)
    sg = torch.sigmoid(pre)
    bb = bias.double().view(1, -1, 1, 1, 1)
    epis = {X.EPI_RELU: F.relu(base + bb), X.EPI_DRELU: base * (pre > 0), X.EPI_DSWISH: base * (sg * (1 + pre * (1 - sg))),
            X.EPI_ADD_AUX: base + auxd, X.EPI_SIGMOID: torch.sigmoid(base + bb), X.EPI_NONE: base + bb}
    for mode, ref in epis.items():
        y = X.new_act(B, N, T, H, W, "cuda")
        stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.float64)
        use_bias = mode in (X.EPI_RELU, X.EPI_SIGMOID, X.EPI_NONE)
        has_aux = mode in (X.EPI_DRELU, X.EPI_DSWISH, X.EPI_ADD_AUX)
        X.pw_conv(rows(x), cu(w), y, B, K, N, g, bias=cu(bias) if use_bias else None, epi=mode, aux=rows(aux) if has_aux else None,
                  epi_tabs=(cu(ea), cu(eb)), stats=stats, stats_mode=X.STATS_SUM_AUX if has_aux else X.STATS_SUM_SQ, tc=True)
        assert relerr(y, ref) <= 5e-6, f"epi {mode}"
        second = (ref * auxd) if has_aux else ref ** 2
        assert relerr(stats[..., 1], second.sum(dim=(2, 3, 4))) <= 2e-5, f"epi {mode} stats"
    y = rows(aux).clone()
    X.pw_conv(rows(x), cu(w), y, B, K, N, g, accumulate=1, tc=True)
    assert relerr(y, base + auxd) <= 5e-6, "accumulate"


def test_tc_large_rows_layer1_shape(X):
    """The bench's dominant launch shape at reduced T: 24 -> 54 channels, 112x112, many row tiles per sample."""
    B, K, N, T, H, W = 2, 24, 54, 3, 112, 112
    x = synth_tensor((B, K, T, H, W), 5)
    w = synth_tensor((N, K), 6, 0.2)
    g = X.geom(T, H, W)
    y1, y2 = X.new_act(B, N, T, H, W, "cuda"), X.new_act(B, N, T, H, W, "cuda")
    s1 = torch.zeros(B, N, 2, device="cuda", dtype=torch.float64)
    s2 = torch.zeros_like(s1)
    X.pw_conv(rows(x), w.cuda(), y1, B, K, N, g, stats=s1, stats_mode=X.STATS_SUM_SQ, tc=True)
    X.pw_conv(rows(x), w.cuda(), y2, B, K, N, g, stats=s2, stats_mode=X.STATS_SUM_SQ, tc=False)
    assert (y1 - y2).abs().max().item() <= 2e-5 * y2.abs().max().item()
    n = T * H * W
    scale = (s2[..., 1] * n).sqrt().unsqueeze(-1)          # sqrt(n * sum y^2) >= sum |y|: the scale of the summands
    assert ((s1 - s2).abs() / torch.stack([scale[..., 0], s2[..., 1]], -1)).max().item() <= 2e-6


@pytest.mark.parametrize("K,N", [(24, 54), (54, 24), (48, 108), (108, 48), (96, 216), (216, 96), (192, 432), (432, 192),
                                 (360, 24), (34, 130), (8, 16)])
def test_tc_wgrad_vs_fp64(X, K, N):
    """Tensor-core weight gradient (MN-major operands, partial kept in TMEM) against fp64, all prologue modes,
    ragged row blocks (R = 4 * 29 * 19 = 2204 rows per sample is not a multiple of 16 or 32), TMEM splits
    (432 x 192 and 192 x 432 exceed 512 columns)."""
    B, T, H, W = 3, 4, 29, 19
    dy, dy2 = synth_tensor((B, N, T, H, W), 21), synth_tensor((B, N, T, H, W), 22)
    x = synth_tensor((B, K, T, H, W), 23)
    da, db, dc = synth_tensor((B, N), 24), synth_tensor((B, N), 25), synth_tensor((B, N), 26)
    ta, tb = synth_tensor((B, K), 27), synth_tensor((B, K), 28)
    v = lambda t: t.double().view(B, -1, 1, 1, 1)
    g = X.geom(T, H, W)
    cu = lambda t: t.cuda()
    # plain
    dw = torch.zeros(N, K, device="cuda")
    X.pw_wgrad(rows(dy), rows(x), dw, B, K, N, g)
    ref = torch.einsum("bnthw,bkthw->nk", dy.double(), x.double())
    assert relerr(dw, ref) <= 1e-5, "plain"
    # accumulation into a non-zero buffer (+=)
    X.pw_wgrad(rows(dy), rows(x), dw, B, K, N, g)
    assert relerr(dw, 2 * ref) <= 1e-5, "accumulate"
    # BatchNorm-backward map on dy, Swish / ReLU prologue on x
    dyy = v(da) * dy.double() + v(db) * dy2.double() + v(dc)
    z = v(ta) * x.double() + v(tb)
    for mode, xx in ((X.PRO_AFFINE_SWISH, z * torch.sigmoid(z)), (X.PRO_AFFINE_RELU, F.relu(z)), (X.PRO_NONE, x.double())):
        dw = torch.zeros(N, K, device="cuda")
        X.pw_wgrad(rows(dy), rows(x), dw, B, K, N, g, dy2=rows(dy2), dy_mode=X.PRO_AFFINE2, dy_tabs=(cu(da), cu(db), cu(dc)),
                   x_mode=mode, x_tabs=(cu(ta), cu(tb)))
        assert relerr(dw, torch.einsum("bnthw,bkthw->nk", dyy, xx)) <= 1e-5, f"x_mode {mode}"


@pytest.mark.parametrize("K,N", [(432, 432), (96, 96), (192, 24), (34, 130)])
def test_tc_wgrad_with_bias_gradient(X, K, N):
    """Weight gradient with a bias gradient (the k=1 Conv1d layers of the fusion block, x3d_coarse.py:216-219): the GEMM
    half runs on the tensor-core kernel and the bias half as a column sum; both accumulate (+=).  R*B = 6612 rows is
    above the tensor-core row threshold and not a multiple of any tile size."""
    B, T, H, W = 3, 4, 29, 19
    dy, x = synth_tensor((B, N, T, H, W), 31), synth_tensor((B, K, T, H, W), 32)
    g = X.geom(T, H, W)
    dw = torch.zeros(N, K, device="cuda")
    db = torch.ones(N, device="cuda")
    n0 = X.lib.cf_launch_count()
    X.pw_wgrad(rows(dy), rows(x), dw, B, K, N, g, dbias=db)
    assert X.lib.cf_launch_count() - n0 == 2, "expected tensor-core GEMM + bias column sum"
    assert relerr(dw, torch.einsum("bnthw,bkthw->nk", dy.double(), x.double())) <= 1e-5
    assert relerr(db, 1.0 + dy.double().sum(dim=(0, 2, 3, 4))) <= 1e-5


@pytest.mark.parametrize("Ti,Hi,Wi", [(9, 11, 14), (8, 14, 14), (4, 7, 7), (1, 2, 5)])
@pytest.mark.parametrize("affine2", [False, True])
def test_dense_s2_dgrad_gather_form(X, Ti, Hi, Wi, affine2):
    """Data gradient of the dense 3x3x3 stride-2 pad-1 conv (pool_1.conv1/conv2, x3d_coarse.py:362-365) through cf_pw_conv
    with scatter_out: 24 channels take the gather-form kernel (one launch, every input position written once, so the
    output buffer may hold garbage); checked against autograd of F.conv3d in fp64, odd and even extents, with and without
    the BatchNorm-backward prologue P*dz + Q*y + R."""
    B, C = 2, 24
    o = lambda n: (n - 1) // 2 + 1
    To, Ho, Wo = o(Ti), o(Hi), o(Wi)
    w = synth_tensor((C, C, 3, 3, 3), 41, 0.1)
    dz, y = synth_tensor((B, C, To, Ho, Wo), 42), synth_tensor((B, C, To, Ho, Wo), 43)
    P, Q, R = synth_tensor((B, C), 44), synth_tensor((B, C), 45), synth_tensor((B, C), 46)
    v = lambda t: t.double().view(B, C, 1, 1, 1)
    gout = v(P) * dz.double() + v(Q) * y.double() + v(R) if affine2 else dz.double()
    xin = torch.zeros(B, C, Ti, Hi, Wi, dtype=torch.float64, requires_grad=True)
    F.conv3d(xin, w.double(), stride=2, padding=1).backward(gout)
    g = X.geom(To, Ho, Wo, Ti, Hi, Wi, k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1), pos_stride=C, sample_stride=Ti * Hi * Wi * C)
    dx = torch.full((B, C, Ti, Hi, Wi), float("nan"), device="cuda").contiguous(memory_format=CL3)
    n0 = X.lib.cf_launch_count()
    kw = dict(x2=rows(y), pro=X.PRO_AFFINE2, pro_tabs=(P.cuda(), Q.cuda(), R.cuda())) if affine2 else {}
    X.pw_conv(rows(dz), w.reshape(C, C * 27).cuda(), dx, B, C, C * 27, g, w_sn=1, w_sk=C * 27, scatter_out=1, **kw)
    assert X.lib.cf_launch_count() - n0 == 1
    assert relerr(dx, xin.grad) <= 2e-6


@pytest.mark.parametrize("Ti,Hi,Wi", [(9, 11, 14), (8, 14, 14), (4, 7, 7), (1, 2, 5), (16, 28, 28)])
@pytest.mark.parametrize("pro", ["none", "affine_relu"])
def test_dense_s2_forward_direct(X, Ti, Hi, Wi, pro):
    """Forward of the dense 3x3x3 stride-2 pad-1 24->24 conv with bias (pool_1.conv1/conv2) through cf_pw_conv with
    gather_in: the direct kernel (one launch) against F.conv3d in fp64, with the bn+ReLU prologue of conv2 (zero padding
    AFTER the prologue) and the BatchNorm statistics of the output; (16,28,28) spans several CTAs per sample."""
    B, C = 2, 24
    o = lambda n: (n - 1) // 2 + 1
    To, Ho, Wo = o(Ti), o(Hi), o(Wi)
    w, bias = synth_tensor((C, C, 3, 3, 3), 51, 0.1), synth_tensor((C,), 52)
    x = synth_tensor((B, C, Ti, Hi, Wi), 53)
    pa, pb = synth_tensor((B, C), 54), synth_tensor((B, C), 55)
    xin = x.double()
    if pro == "affine_relu":
        xin = F.relu(pa.double().view(B, C, 1, 1, 1) * xin + pb.double().view(B, C, 1, 1, 1))
    ref = F.conv3d(xin, w.double(), bias.double(), stride=2, padding=1)
    g = X.geom(To, Ho, Wo, Ti, Hi, Wi, k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1), pos_stride=C, sample_stride=Ti * Hi * Wi * C)
    y = torch.full((B, C, To, Ho, Wo), float("nan"), device="cuda").contiguous(memory_format=CL3)
    stats = torch.zeros(B, C, 2, device="cuda", dtype=torch.float64)
    kw = dict(pro=X.PRO_AFFINE_RELU, pro_tabs=(pa.cuda(), pb.cuda(), None)) if pro == "affine_relu" else {}
    n0 = X.lib.cf_launch_count()
    X.pw_conv(rows(x), w.reshape(C, C * 27).cuda(), y, B, C * 27, C, g, bias=bias.cuda(), gather_in=1, stats=stats,
              stats_mode=X.STATS_SUM_SQ, **kw)
    assert X.lib.cf_launch_count() - n0 == 1
    assert relerr(y, ref) <= 2e-6
    assert relerr(stats[..., 0], ref.sum(dim=(2, 3, 4))) <= 1e-5
    assert relerr(stats[..., 1], (ref * ref).sum(dim=(2, 3, 4))) <= 1e-5


@pytest.mark.parametrize("Ti,Hi,Wi", [(9, 11, 14), (4, 7, 7), (1, 2, 5), (16, 28, 28)])
@pytest.mark.parametrize("modes", ["plain", "bn"])
def test_dense_s2_wgrad_direct(X, Ti, Hi, Wi, modes):
    """Weight + bias gradient of the dense 3x3x3 stride-2 pad-1 24->24 conv (pool_1.conv1/conv2) through cf_pw_wgrad with
    gather_in: one launch of the staged-row kernel, accumulating (+=) into non-zero buffers; against autograd of F.conv3d
    in fp64, with the BatchNorm-backward map on the output gradient and the bn+ReLU prologue on the input ("bn")."""
    B, C = 2, 24
    o = lambda n: (n - 1) // 2 + 1
    To, Ho, Wo = o(Ti), o(Hi), o(Wi)
    x = synth_tensor((B, C, Ti, Hi, Wi), 61)
    dz, y = synth_tensor((B, C, To, Ho, Wo), 62), synth_tensor((B, C, To, Ho, Wo), 63)
    P, Q, R = synth_tensor((B, C), 64), synth_tensor((B, C), 65), synth_tensor((B, C), 66)
    xa, xb = synth_tensor((B, C), 67), synth_tensor((B, C), 68)
    v = lambda t: t.double().view(B, C, 1, 1, 1)
    bn = modes == "bn"
    gout = v(P) * dz.double() + v(Q) * y.double() + v(R) if bn else dz.double()
    xin = F.relu(v(xa) * x.double() + v(xb)) if bn else x.double()
    w = torch.zeros(C, C, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    bias = torch.zeros(C, dtype=torch.float64, requires_grad=True)
    F.conv3d(xin, w, bias, stride=2, padding=1).backward(gout)
    g = X.geom(To, Ho, Wo, Ti, Hi, Wi, k=(3, 3, 3), s=(2, 2, 2), p=(1, 1, 1), pos_stride=C, sample_stride=Ti * Hi * Wi * C)
    dw = torch.ones(C, C * 27, device="cuda")
    db = torch.ones(C, device="cuda")
    kw = dict(dy2=rows(y), dy_mode=X.PRO_AFFINE2, dy_tabs=(P.cuda(), Q.cuda(), R.cuda()), x_mode=X.PRO_AFFINE_RELU,
              x_tabs=(xa.cuda(), xb.cuda())) if bn else {}
    n0 = X.lib.cf_launch_count()
    X.pw_wgrad(rows(dz), rows(x), dw, B, C * 27, C, g, dbias=db, gather_in=1, **kw)
    assert X.lib.cf_launch_count() - n0 == 1
    assert relerr(dw, 1.0 + w.grad.reshape(C, C * 27)) <= 1e-5
    assert relerr(db, 1.0 + bias.grad) <= 1e-5


@pytest.mark.parametrize("K,N,T,H,W", [(54, 24, 1, 7, 7), (54, 24, 3, 7, 7), (24, 54, 1, 5, 9), (108, 48, 5, 7, 7), (6, 10, 2, 9, 8),
                                       (216, 96, 17, 7, 7), (432, 192, 17, 7, 7)])
def test_tc_tma_edge_shapes(X, K, N, T, H, W):
    """Shapes around the TMA-fed producers' eligibility rules: fewer rows than one tile, odd row counts (a 54-channel row
    tensor with an odd number of rows cannot be folded into 16-byte-aligned TMA rows: register-load producers), sample
    boundaries inside a tile-sized box (rows of the next sample must not leak in), k-chunks that end inside a box."""
    B = 3
    x, x2 = synth_tensor((B, K, T, H, W), 31), synth_tensor((B, K, T, H, W), 32)
    w = synth_tensor((N, K), 33, 0.2)
    ta, tb, tcc = synth_tensor((B, K), 34), synth_tensor((B, K), 35), synth_tensor((B, K), 36)
    v = lambda t: t.double().view(B, -1, 1, 1, 1)
    g = X.geom(T, H, W)
    conv = lambda z: torch.einsum("nk,bkthw->bnthw", w.double(), z.double())
    for mode, xin in ((X.PRO_NONE, x.double()), (X.PRO_AFFINE_RELU, F.relu(v(ta) * x.double() + v(tb))),
                      (X.PRO_AFFINE2, v(ta) * x.double() + v(tb) * x2.double() + v(tcc))):
        y = X.new_act(B, N, T, H, W, "cuda")
        stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.float64)
        X.pw_conv(rows(x), w.cuda(), y, B, K, N, g, x2=rows(x2) if mode == X.PRO_AFFINE2 else None, pro=mode,
                  pro_tabs=(ta.cuda(), tb.cuda(), tcc.cuda()) if mode != X.PRO_NONE else (None, None, None), stats=stats,
                  stats_mode=X.STATS_SUM_SQ, tc=True)
        ref = conv(xin)
        assert relerr(y, ref) <= 1e-5, f"pro {mode}"
        assert relerr(stats[..., 0], ref.sum(dim=(2, 3, 4))) <= 1e-4 and relerr(stats[..., 1], (ref ** 2).sum(dim=(2, 3, 4))) <= 3e-5
