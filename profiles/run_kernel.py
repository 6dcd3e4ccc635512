"""Launch one hot kernel of the step in isolation (for `ncu --set full`); shapes = the bench's largest launch.

    python profiles/run_kernel.py pw_tc | pw_simt | pw_swish | pw_dgrad3 | pw_dgrad1 | wgrad1 | wgrad3 | dw_dgrad | dw_wgrad | dw_fwd | dw_fused [B] [T]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
from coarse_fine_networks_b200 import x3d_ops as X  # noqa: E402

which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda")
CL3 = torch.channels_last_3d
if which in ("pw_tc", "pw_simt"):
    K, N, H, W = 24, 54, 112, 112
    x = torch.randn(B, K, T, H, W, device=dev).contiguous(memory_format=CL3)
    w = torch.randn(N, K, device=dev) * 0.1
    y = X.new_act(B, N, T, H, W, dev)
    stats = torch.zeros(B, N, 2, device=dev, dtype=torch.float64)
    fn = lambda: X.pw_conv(x, w, y, B, K, N, X.geom(T, H, W), stats=stats, stats_mode=X.STATS_SUM_SQ, tc=which == "pw_tc")
elif which in ("pw_swish", "pw_dgrad3", "pw_dgrad1"):
    # layer-1 shapes at 56x56: conv3 forward (54 -> 24, BN+SE+Swish prologue), conv3 data gradient (24 -> 54, BN-backward
    # prologue, Swish' epilogue with aux, statistics), conv1 data gradient (54 -> 24, BN-backward prologue, residual add)
    K, N = {"pw_swish": (54, 24), "pw_dgrad3": (24, 54), "pw_dgrad1": (54, 24)}[which]
    H = W = 56
    pro = X.PRO_AFFINE_SWISH if which == "pw_swish" else X.PRO_AFFINE2
    epi = {"pw_swish": X.EPI_NONE, "pw_dgrad3": X.EPI_DSWISH, "pw_dgrad1": X.EPI_ADD_AUX}[which]
    smode = {"pw_swish": X.STATS_SUM_SQ, "pw_dgrad3": X.STATS_SUM_AUX, "pw_dgrad1": X.STATS_NONE}[which]
    x = torch.randn(B, K, T, H, W, device=dev).contiguous(memory_format=CL3)
    x2 = torch.randn_like(x) if pro == X.PRO_AFFINE2 else None
    w = torch.randn(N, K, device=dev) * 0.1
    y = X.new_act(B, N, T, H, W, dev)
    aux = torch.randn_like(y) if which != "pw_swish" else None
    tabs = tuple(torch.randn(B, K, device=dev) for _ in range(3))
    etabs = (torch.randn(B, N, device=dev), torch.randn(B, N, device=dev)) if epi == X.EPI_DSWISH else (None, None)
    stats = torch.zeros(B, N, 2, device=dev, dtype=torch.float64) if smode != X.STATS_NONE else None
    fn = lambda: X.pw_conv(x, w, y, B, K, N, X.geom(T, H, W), x2=x2, pro=pro, pro_tabs=tabs, epi=epi, aux=aux, epi_tabs=etabs,
                           stats=stats, stats_mode=smode, tc=True)
elif which in ("wgrad1", "wgrad3"):
    K, N, xmode = (24, 54, X.PRO_NONE) if which == "wgrad1" else (54, 24, X.PRO_AFFINE_SWISH)
    H = W = 56
    dy = torch.randn(B, N, T, H, W, device=dev).contiguous(memory_format=CL3)
    dy2 = torch.randn_like(dy)
    x = torch.randn(B, K, T, H, W, device=dev).contiguous(memory_format=CL3)
    dtabs = tuple(torch.randn(B, N, device=dev) for _ in range(3))
    xtabs = (torch.randn(B, K, device=dev), torch.randn(B, K, device=dev)) if xmode != X.PRO_NONE else (None, None)
    dw = torch.zeros(N, K, device=dev)
    fn = lambda: X.pw_wgrad(dy, x, dw, B, K, N, X.geom(T, H, W), dy2=dy2, dy_mode=X.PRO_AFFINE2, dy_tabs=dtabs, x_mode=xmode, x_tabs=xtabs)
elif which == "dw_fused":
    C, H, W = 54, 56, 56                                   # stride-1 depthwise: data gradient + weight gradient in one pass
    g = X.geom(T, H, W, k=(3, 3, 3), p=(1, 1, 1))
    y1 = torch.randn(B, C, T, H, W, device=dev).contiguous(memory_format=CL3)
    y2, dU = torch.randn_like(y1), torch.randn_like(y1)
    w = torch.randn(C, 27, device=dev) * 0.1
    tabs = [torch.randn(B, C, device=dev) for _ in range(5)]
    stats = torch.zeros(B, C, 2, device=dev, dtype=torch.float64)
    dz1, dw = torch.empty_like(y1), torch.zeros(C, 27, device=dev)
    fn = lambda: X.dw_call("cf_dw_conv_dgrad", dU, w, dz1, B, C, g, x2=y2, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs[:3]), aux=y1,
                           epi=X.EPI_DRELU, epi_tabs=(tabs[3], tabs[4]), stats=stats, stats_mode=X.STATS_SUM_AUX, dw_out=dw)
else:
    C, H, W, s = 54, 112, 112, 2
    Ho, Wo = H // s, W // s
    g = X.geom(T, Ho, Wo, T, H, W, k=(3, 3, 3), s=(1, s, s), p=(1, 1, 1))
    y1 = torch.randn(B, C, T, H, W, device=dev).contiguous(memory_format=CL3)
    y2 = torch.randn(B, C, T, Ho, Wo, device=dev).contiguous(memory_format=CL3)
    dU = torch.randn_like(y2)
    w = torch.randn(C, 27, device=dev) * 0.1
    tabs = [torch.randn(B, C, device=dev) for _ in range(5)]
    stats = torch.zeros(B, C, 2, device=dev, dtype=torch.float64)
    if which == "dw_fwd":
        out = torch.empty_like(y2)
        fn = lambda: X.dw_call("cf_dw_conv_fwd", y1, w, out, B, C, g, pro=X.PRO_AFFINE_RELU, pro_tabs=(tabs[0], tabs[1], None),
                               stats=stats, stats_mode=X.STATS_SUM_SQ)
    elif which == "dw_dgrad":
        dz1 = torch.empty_like(y1)
        fn = lambda: X.dw_call("cf_dw_conv_dgrad", dU, w, dz1, B, C, g, x2=y2, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs[:3]), aux=y1,
                               epi=X.EPI_DRELU, epi_tabs=(tabs[3], tabs[4]), stats=stats, stats_mode=X.STATS_SUM_AUX)
    else:
        dw = torch.zeros(C, 27, device=dev)
        fn = lambda: X.dw_call("cf_dw_conv_wgrad", dU, w, dw, B, C, g, x2=y2, pro=X.PRO_AFFINE2, pro_tabs=tuple(tabs[:3]), aux=y1,
                               epi_tabs=(tabs[3], tabs[4]))
for _ in range(3):
    fn()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(5):
    fn()
ev[1].record()
torch.cuda.synchronize()
print(which, "B", B, "T", T, "ms per launch", ev[0].elapsed_time(ev[1]) / 5)
