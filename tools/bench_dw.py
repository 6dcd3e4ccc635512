"""Time the depthwise 3x3x3 kernels (forward, data gradient, weight gradient) on the fine stream's stage shapes.

    python tools/bench_dw.py            # plane-marching kernels (x3d_dw3.cu) where eligible
    CFNET_DW3_OFF=1 python tools/bench_dw.py
Algorithmic bytes: fwd read x + write y; dgrad read dU, y2, y1 + write dz; wgrad read dU, y2, y1.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
from coarse_fine_networks_b200 import x3d_ops as X  # noqa: E402

dev = torch.device("cuda")
CL3 = torch.channels_last_3d
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
peak = 6531.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=6):
    for _ in range(2):
        fn()
    ev = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


B = 4
CASES = [("layer1 54ch 56x56", 54, 256, 56, 56, 1), ("layer2 108ch 28x28", 108, 256, 28, 28, 1), ("layer3 216ch 14x14", 216, 256, 14, 14, 1),
         ("layer4 432ch 7x7", 432, 256, 7, 7, 1), ("layer1.0 54ch 112->56 s2", 54, 64, 112, 112, 2)]
only = sys.argv[1] if len(sys.argv) > 1 else None
print("depthwise kernels:", "general (x3d_dw.cu)" if os.environ.get("CFNET_DW3_OFF") == "1" else "plane-marching (x3d_dw3.cu) where eligible")
for name, C, T, H, W, s in CASES:
    if only and only not in name:
        continue
    Ho, Wo = H // s, W // s
    g = X.geom(T, Ho, Wo, T, H, W, k=(3, 3, 3), s=(1, s, s), p=(1, 1, 1))
    y1 = torch.randn(B, C, T, H, W, device=dev).contiguous(memory_format=CL3)
    y2 = torch.randn(B, C, T, Ho, Wo, device=dev).contiguous(memory_format=CL3)
    dU = torch.randn_like(y2)
    w = torch.randn(C, 27, device=dev) * 0.1
    tabs = [torch.randn(B, C, device=dev) for _ in range(5)]
    stats = torch.zeros(B, C, 2, device=dev, dtype=torch.float64)
    out = torch.empty_like(y2)
    dz1 = torch.empty_like(y1)
    dw = torch.zeros(C, 27, device=dev)
    n_in, n_out = y1.numel(), y2.numel()
    runs = [
        ("fwd", lambda: X.dw_call("cf_dw_conv_fwd", y1, w, out, B, C, g, pro=X.PRO_AFFINE_RELU, pro_tabs=(tabs[0], tabs[1], None),
                                  stats=stats, stats_mode=X.STATS_SUM_SQ), 4 * (n_in + n_out)),
        ("dgrad", lambda: X.dw_call("cf_dw_conv_dgrad", dU, w, dz1, B, C, g, x2=y2, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs[:3]), aux=y1,
                                    epi=X.EPI_DRELU, epi_tabs=(tabs[3], tabs[4]), stats=stats, stats_mode=X.STATS_SUM_AUX),
         4 * (2 * n_out + 2 * n_in)),
        ("wgrad", lambda: X.dw_call("cf_dw_conv_wgrad", dU, w, dw, B, C, g, x2=y2, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs[:3]), aux=y1,
                                    epi_tabs=(tabs[3], tabs[4])), 4 * (2 * n_out + n_in)),
        ("fused", lambda: X.dw_call("cf_dw_conv_dgrad", dU, w, dz1, B, C, g, x2=y2, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs[:3]), aux=y1,
                                    epi=X.EPI_DRELU, epi_tabs=(tabs[3], tabs[4]), stats=stats, stats_mode=X.STATS_SUM_AUX, dw_out=dw),
         4 * (2 * n_out + 2 * n_in)),
    ]
    for kind, fn, byt in runs:
        ms = timeit(fn)
        print(f"{name:28s} {kind:6s} {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s  {100*byt/ms/1e6/peak:5.1f}% of measured HBM peak")
    del y1, y2, dU, out, dz1

# ---- stem temporal 5x1x1 depthwise (conv1_t): 24 channels at 112x112
if not only or "stem" in only:
    C, T, H, W = 24, 256, 112, 112
    g = X.geom(T, H, W, k=(5, 1, 1), p=(2, 0, 0))
    y0 = torch.randn(B, C, T, H, W, device=dev).contiguous(memory_format=CL3)
    yt = torch.randn_like(y0)
    dz = torch.randn_like(y0)
    out = torch.empty_like(y0)
    w = torch.randn(C, 5, device=dev) * 0.3
    tabs = [torch.randn(B, C, device=dev) for _ in range(3)]
    stats = torch.zeros(B, C, 2, device=dev, dtype=torch.float64)
    dw = torch.zeros(C, 5, device=dev)
    n = y0.numel()
    for kind, fn, byt in [
        ("fwd", lambda: X.dw_call("cf_dw_conv_fwd", y0, w, out, B, C, g, stats=stats, stats_mode=X.STATS_SUM_SQ), 8 * n),
        ("dgrad", lambda: X.dw_call("cf_dw_conv_dgrad", dz, w, out, B, C, g, x2=yt, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs)), 12 * n),
        ("wgrad", lambda: X.dw_call("cf_dw_conv_wgrad", dz, w, dw, B, C, g, x2=yt, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs), aux=y0), 12 * n)]:
        ms = timeit(fn)
        print(f"{'stem conv1_t 24ch 112x112':28s} {kind:6s} {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s  {100*byt/ms/1e6/peak:5.1f}% of measured HBM peak")
