"""GPU parity: Grid Pool / Grid Unpool CUDA kernels (through the C ABI) vs the oracle and the
golden vectors produced by the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import cf_oracle as O
from synth import synth_tensor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def close(a, b, rtol=1e-4, atol=1e-5):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= atol + rtol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.fixture(scope="module")
def ops():
    from coarse_fine_networks_b200 import gridpool_ops
    return gridpool_ops


def test_cdf_fwd_bwd(ops):
    for name in ("gridpool_layer", "gridpool_t64"):
        g = load(name)
        cdf = ops.gridpool_cdf(g["g"].cuda())
        close(cdf, g["cdf"], rtol=0, atol=2e-6)
    gg = (synth_tensor((37, 16), 101) * 2).requires_grad_(True)
    w = synth_tensor((37, 17), 102)
    (O.gridpool_cdf(gg) * w).sum().backward()
    gc = gg.detach().cuda().requires_grad_(True)
    (ops.gridpool_cdf(gc) * w.cuda()).sum().backward()
    close(gc.grad, gg.grad, rtol=1e-4, atol=1e-7)
    # long rows exercise the chunked scan (n > 32)
    g2 = synth_tensor((5, 100), 103)
    close(ops.gridpool_cdf(g2.cuda()), O.gridpool_cdf(g2), rtol=0, atol=3e-6)


def test_bins_bit_exact_given_cdf(ops):
    for name, t in (("gridpool_layer", 16), ("gridpool_t64", 64)):
        cdf = load(name)["cdf"]
        i0, w1 = ops.sample_bins(cdf.cuda(), t)
        _, ri0, rw1 = O.sample_coords(cdf, t)
        assert torch.equal(i0.cpu().long(), ri0)
        assert torch.equal(w1.cpu(), rw1)
    # stress: random monotone cdfs incl. exact 0 / 1 endpoints and many T
    g = torch.Generator().manual_seed(5)
    for t in (8, 13, 64, 256):
        c = torch.sort(torch.rand(64, 33, generator=g), dim=1)[0]
        c[:, 0] = 0
        c[::2, -1] = 1
        i0, w1 = ops.sample_bins(c.cuda(), t)
        _, ri0, rw1 = O.sample_coords(c, t)
        assert torch.equal(i0.cpu().long(), ri0)
        assert torch.equal(w1.cpu(), rw1)


@pytest.mark.parametrize("channels_last", [False, True])
def test_gather_golden(ops, channels_last):
    for name in ("gridpool_layer", "gridpool_t64"):
        g = load(name)
        x = g["x"].cuda()
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last_3d)
        out = ops.temporal_sample(x, g["cdf"].cuda())
        assert out.shape == g["out"].shape
        close(out, g["out"], rtol=2e-5, atol=1e-6)          # vs the reference's 5-D grid_sample


@pytest.mark.parametrize("channels_last", [False, True])
def test_gather_backward_golden(ops, channels_last):
    g = load("gridpool_layer")
    x = g["x"].cuda()
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last_3d)
    x.requires_grad_(True)
    cdf = g["cdf"].cuda().requires_grad_(True)
    out = ops.temporal_sample(x, cdf)
    (out * g["gout"].cuda()).sum().backward()
    close(x.grad, g["gather_dx"], rtol=2e-5, atol=1e-6)
    close(cdf.grad, g["gather_dcdf"], rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 3, 9, 5, 7), (1, 5, 32, 4, 4), (3, 2, 7, 1, 1), (2, 4, 12, 6, 2)])
def test_gather_ragged_shapes_vs_oracle(ops, shape):
    """odd inner sizes take the scalar path; random (non-monotone, out-of-range) coordinates
    exercise zero padding."""
    x = synth_tensor(shape, 201).requires_grad_(True)
    coord = (torch.rand(shape[0], 11, generator=torch.Generator().manual_seed(7)) * 1.4 - 0.2).requires_grad_(True)
    gout = synth_tensor((shape[0], shape[1], 11) + shape[3:], 202)
    (O.temporal_lerp(x, coord) * gout).sum().backward()
    xc = x.detach().cuda().requires_grad_(True)
    cc = coord.detach().cuda().requires_grad_(True)
    out = ops.temporal_sample(xc, cc)
    close(out, O.temporal_lerp(x, coord), rtol=1e-5, atol=1e-6)
    (out * gout.cuda()).sum().backward()
    close(xc.grad, x.grad, rtol=1e-5, atol=1e-6)
    close(cc.grad, coord.grad, rtol=1e-4, atol=1e-4)


def test_inverse_cdf_bit_exact(ops):
    g = load("interp1d")
    inv, ind = ops.inverse_cdf(g["x"].cuda())
    assert torch.equal(ind.cpu().long(), g["ind"])
    assert torch.equal(inv.cpu(), g["ynew"])
    g2 = load("gridunpool")
    inv2, ind2 = ops.inverse_cdf(g2["cdf"].cuda())
    rinv, rind = O.inverse_cdf(g2["cdf"])
    assert torch.equal(ind2.cpu().long(), rind)
    assert torch.equal(inv2.cpu(), rinv)


def test_gridunpool_golden(ops):
    g = load("gridunpool")
    x = g["x"].cuda().requires_grad_(True)
    cdf = g["cdf"].cuda().requires_grad_(True)
    y = ops.grid_unpool(x, cdf, True)
    close(y, g["y"], rtol=1e-5, atol=1e-6)
    y_up = ops.linear_upsample_t(y, (y.shape[2] - 1) * 4)
    close(y_up, g["y_up"], rtol=1e-5, atol=1e-6)
    (y_up * g["gout"].cuda()).sum().backward()
    close(x.grad, g["dx"], rtol=1e-5, atol=1e-6)
    close(cdf.grad, g["dcdf"], rtol=2e-4, atol=1e-4)
    for cl in (False, True):
        xf = g["xf"].cuda()
        if cl:
            xf = xf.contiguous(memory_format=torch.channels_last_3d)
        yf = ops.grid_unpool(xf, g["cdf"].cuda(), False)
        close(yf, g["yf"], rtol=1e-5, atol=1e-6)


def test_cfg3_full_size_properties(ops):
    """BASELINE cfg 3: [32,24,64,56,56].  Checked against the oracle formula evaluated by
    torch on the GPU (checker only), plus linearity and end-point properties."""
    torch.manual_seed(0)
    B, C, T, H, W = 32, 24, 64, 56, 56
    x = torch.randn(B, C, T, H, W, device="cuda")
    cdf = ops.gridpool_cdf(torch.randn(B, 16, device="cuda") * 2)
    out = ops.temporal_sample(x, cdf)
    ref = O.temporal_lerp(x, cdf)
    assert out.shape == (B, C, 17, H, W)
    assert (out - ref).abs().max().item() < 1e-5
    assert torch.equal(out[:, :, 0], x[:, :, 0])                       # cdf[0] = 0 -> frame 0 exactly
    y = torch.randn_like(x)
    lin = ops.temporal_sample(2.0 * x + y, cdf) - (2.0 * out + ops.temporal_sample(y, cdf))
    assert lin.abs().max().item() < 1e-4
    del y, lin, ref
    # adjoint property <gather(x), g> == <x, gather^T(g)>
    gout = torch.randn_like(out)
    xr = x.requires_grad_(True)
    o2 = ops.temporal_sample(xr, cdf)
    (o2 * gout).sum().backward()
    lhs = (o2.double() * gout.double()).sum().item()
    rhs = (xr.grad.double() * x.detach().double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0) + 1e-3
    # frames no sample point touches receive exactly zero gradient
    i0, w1 = ops.sample_bins(cdf, T)
    touched = torch.zeros(B, T + 1, device="cuda", dtype=torch.bool)
    touched.scatter_(1, i0.long().clamp(0, T), True)
    touched.scatter_(1, (i0.long() + 1).clamp(0, T), True)
    untouched = ~touched[:, :T]
    gsum = xr.grad.abs().sum(dim=(1, 3, 4))
    assert float(gsum[untouched].abs().max()) == 0.0
