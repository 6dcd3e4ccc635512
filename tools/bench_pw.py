"""Time the pointwise-conv GEMM on the step's representative launches (CUDA events, L2 flushed between launches).

    python tools/bench_pw.py            # persistent kernel
    CFNET_PW_TC_V1=1 python tools/bench_pw.py
Prints algorithmic GB/s (read x [+x2] [+aux] + write y) per shape.
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


def timeit(fn, reps=8):
    for _ in range(3):
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


# (name, K, N, T, H, W, pro, epi, stats)
B = 4
CASES = [
    ("l1.0 conv1 fwd 24->54 @112", 24, 54, 64, 112, 112, X.PRO_NONE, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l1 conv1 fwd 24->54 @56", 24, 54, 256, 56, 56, X.PRO_NONE, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l1 conv3 fwd 54->24 @56 swish", 54, 24, 256, 56, 56, X.PRO_AFFINE_SWISH, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l1 conv3 dgrad 24->54 aff2+dswish", 24, 54, 256, 56, 56, X.PRO_AFFINE2, X.EPI_DSWISH, X.STATS_SUM_AUX),
    ("l1 conv1 dgrad 54->24 aff2+add", 54, 24, 256, 56, 56, X.PRO_AFFINE2, X.EPI_ADD_AUX, X.STATS_NONE),
    ("l2 conv1 fwd 48->108 @28", 48, 108, 256, 28, 28, X.PRO_NONE, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l2 conv3 fwd 108->48 @28 swish", 108, 48, 256, 28, 28, X.PRO_AFFINE_SWISH, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l3 conv1 fwd 96->216 @14", 96, 216, 256, 14, 14, X.PRO_NONE, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l3 conv3 fwd 216->96 @14 swish", 216, 96, 256, 14, 14, X.PRO_AFFINE_SWISH, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l3 conv3 dgrad 96->216 aff2+dswish", 96, 216, 256, 14, 14, X.PRO_AFFINE2, X.EPI_DSWISH, X.STATS_SUM_AUX),
    ("l4 conv1 fwd 192->432 @7", 192, 432, 256, 7, 7, X.PRO_NONE, X.EPI_NONE, X.STATS_SUM_SQ),
    ("l4 conv3 fwd 432->192 @7 swish", 432, 192, 256, 7, 7, X.PRO_AFFINE_SWISH, X.EPI_NONE, X.STATS_SUM_SQ),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
print("kernel:", "v1 (one tile per CTA)" if os.environ.get("CFNET_PW_TC_V1") == "1" else "persistent warp-specialised")
for name, K, N, T, H, W, pro, epi, smode in CASES:
    if only and only not in name:
        continue
    x = torch.randn(B, K, T, H, W, device=dev).contiguous(memory_format=CL3)
    x2 = torch.randn_like(x) if pro == X.PRO_AFFINE2 else None
    w = torch.randn(N, K, device=dev) * 0.1
    y = X.new_act(B, N, T, H, W, dev)
    need_aux = epi in (X.EPI_DRELU, X.EPI_DSWISH, X.EPI_ADD_AUX) or smode == X.STATS_SUM_AUX
    aux = torch.randn_like(y) if need_aux else None
    tabs = tuple(torch.randn(B, K, device=dev) for _ in range(3)) if pro != X.PRO_NONE else (None, None, None)
    etabs = (torch.randn(B, N, device=dev), torch.randn(B, N, device=dev)) if epi in (X.EPI_DRELU, X.EPI_DSWISH) else (None, None)
    stats = torch.zeros(B, N, 2, device=dev, dtype=torch.float64) if smode != X.STATS_NONE else None
    g = X.geom(T, H, W)
    fn = lambda: X.pw_conv(x, w, y, B, K, N, g, x2=x2, pro=pro, pro_tabs=tabs, epi=epi, aux=aux, epi_tabs=etabs, stats=stats,
                           stats_mode=smode, tc=True)
    ms = timeit(fn)
    rows = B * T * H * W
    byt = rows * 4 * (K * (2 if x2 is not None else 1) + N * (2 if aux is not None else 1))
    if os.environ.get("CFNET_PW_TC_TIMING") == "1":
        import ctypes
        from coarse_fine_networks_b200 import _lib
        buf = (ctypes.c_longlong * 24)()
        _lib.lib.cf_pw_tc_debug_read(buf)          # discard the accumulated launches, then time exactly one
        fn()
        torch.cuda.synchronize()
        _lib.lib.cf_pw_tc_debug_read(buf)
        v = list(buf)
        print("   producer thread 0  [loads, wait-stage, transform+store, fence, arrive] =", v[0:5], "total", v[7])
        print("   mma warp           [bookkeeping, wait-acc, wait-stage, issue+commit] =", v[16:20], "total", v[20])
        print("   epilogue thread 0  [bookkeeping, wait-acc, tmem-ld, barriers, sts, store-slab, stats-reduce] =", v[8:15], "total", v[15])
    print(f"{name:38s} rows {rows:9d}  {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s  {100*byt/ms/1e6/peak:5.1f}% of measured HBM peak")
    del x, x2, y, aux

# ---- weight gradients (dy [rows,N], x [rows,K] -> dw [N,K]); bytes = read dy, dy2, x
WCASES = [
    ("l1 conv1 wgrad dy54 x24", 24, 54, 256, 56, 56, X.PRO_NONE),
    ("l1 conv3 wgrad dy24 x54 swish", 54, 24, 256, 56, 56, X.PRO_AFFINE_SWISH),
    ("l2 conv3 wgrad dy48 x108 swish", 108, 48, 256, 28, 28, X.PRO_AFFINE_SWISH),
    ("l3 conv1 wgrad dy216 x96", 96, 216, 256, 14, 14, X.PRO_NONE),
    ("l3 conv3 wgrad dy96 x216 swish", 216, 96, 256, 14, 14, X.PRO_AFFINE_SWISH),
    ("l4 conv3 wgrad dy192 x432 swish", 432, 192, 256, 7, 7, X.PRO_AFFINE_SWISH),
]
print("weight gradient:", "CUDA-core kernel" if os.environ.get("CFNET_PW_WGRAD_SIMT") == "1" else "tensor-core kernel")
for name, K, N, T, H, W, xmode in WCASES:
    if only and only not in name:
        continue
    dy = torch.randn(B, N, T, H, W, device=dev).contiguous(memory_format=CL3)
    dy2 = torch.randn_like(dy)
    x = torch.randn(B, K, T, H, W, device=dev).contiguous(memory_format=CL3)
    dtabs = tuple(torch.randn(B, N, device=dev) for _ in range(3))
    xtabs = (torch.randn(B, K, device=dev), torch.randn(B, K, device=dev)) if xmode != X.PRO_NONE else (None, None)
    dw = torch.zeros(N, K, device=dev)
    g = X.geom(T, H, W)
    fn = lambda: X.pw_wgrad(dy, x, dw, B, K, N, g, dy2=dy2, dy_mode=X.PRO_AFFINE2, dy_tabs=dtabs, x_mode=xmode, x_tabs=xtabs)
    ms = timeit(fn)
    rows = B * T * H * W
    byt = rows * 4 * (2 * N + K)
    print(f"{name:38s} rows {rows:9d}  {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s  {100*byt/ms/1e6/peak:5.1f}% of measured HBM peak")
    del dy, dy2, x
