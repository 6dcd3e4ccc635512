"""Autograd functions of the X3D conv stacks over the C ABI (cf_pw_conv, cf_pw_wgrad,
cf_dw_conv_*, cf_bn_*, cf_se_*, cf_residual_*, cf_block_avgpool_*).

Activations are channels-last fp32 ([B,C,T,H,W] logical shape, torch.channels_last_3d strides);
BatchNorm / ReLU / SE / Swish never materialise: producers accumulate statistics, consumers
apply per-(sample,channel) affine tables on load (see include/cfnet_b200.h)."""
import torch

from ._lib import STRUCTS, call, call_struct, lib, make, ptr, stream_ptr

CL3 = torch.channels_last_3d

PRO_NONE, PRO_AFFINE, PRO_AFFINE_RELU, PRO_AFFINE_SWISH, PRO_AFFINE2 = 0, 1, 2, 3, 4
EPI_NONE, EPI_RELU, EPI_DRELU, EPI_DSWISH, EPI_ADD_AUX, EPI_SIGMOID, EPI_AFFINE, EPI_AFFINE_ADD_RELU = 0, 1, 2, 3, 4, 5, 6, 7
STATS_NONE, STATS_SUM_SQ, STATS_SUM_AUX = 0, 1, 2


def cl(x):
    """Dense channels-last view of a [B,C,T,H,W] tensor (no copy when already so)."""
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous(memory_format=CL3)


def new_act(B, C, T, H, W, device):
    return torch.empty((B, C, T, H, W), device=device, dtype=torch.float32, memory_format=CL3)


def geom(T, H, W, Ti=None, Hi=None, Wi=None, k=(1, 1, 1), s=(1, 1, 1), p=(0, 0, 0), pos_stride=0, ch_stride=1,
         sample_stride=0):
    g = STRUCTS["cf_geom"]()
    g.T, g.H, g.W = T, H, W
    g.Ti, g.Hi, g.Wi = (T if Ti is None else Ti), (H if Hi is None else Hi), (W if Wi is None else Wi)
    g.kt, g.kh, g.kw = k
    g.st, g.sh, g.sw = s
    g.pt, g.ph, g.pw = p
    g.pos_stride, g.ch_stride, g.sample_stride = pos_stride, ch_stride, sample_stride
    return g


# Dense pointwise GEMMs run on the tcgen05 tensor cores (3xTF32); pw_conv(..., tc=False) selects the fp32 CUDA-core kernel
# (used by the tests to cross-check the two paths; not a fallback: both are sm_100a kernels of the library).
USE_TC = True


class PackCache:
    """Packed (hi/lo TF32 split, swizzled) copies of GEMM weights that live across conv calls.

    Without it every tensor-core pw_conv launches a small packing kernel first (272 per training step).  Weights owned by a
    train.FlatTrainer change exactly once per step, so the trainer registers its flat parameter buffer here; a conv whose
    weight lies inside a registered buffer keeps a persistent pack, and FlatTrainer.step() re-packs ALL of them with ONE
    launch (cf_pw_pack_many) right after the SGD update.  An entry is trusted only if (i) the global weights version is
    the one it was packed at and (ii) the tensor's autograd version counter is unchanged (load_state_dict / in-place edits
    bump it) -- otherwise the call packs again.  Weights outside registered buffers are packed per call as before."""

    def __init__(self):
        self.ranges = []           # [(lo, hi, weakref to the flat buffer)]
        self.entries = {}          # (ptr, w_sn, w_sk, K, N) -> dict(buf, nt, gver, tver, w)
        self.gver = 0
        self._items = None
        self._items_n = -1

    def register(self, flat):
        import weakref
        self.ranges.append((flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size(), weakref.ref(flat)))

    def _owned(self, ptr_):
        alive = [r for r in self.ranges if r[2]() is not None]
        if len(alive) != len(self.ranges):                      # a trainer went away: drop its packs
            self.ranges = alive
            self.entries = {k: e for k, e in self.entries.items() if any(lo <= k[0] < hi for lo, hi, _ in alive)}
            self._items_n = -1
        return any(lo <= ptr_ < hi for lo, hi, _ in self.ranges)

    def lookup(self, w, w_sn, w_sk, K, N, nbytes):
        """-> (buffer, wpack_nt) for a weight inside a registered buffer, else None."""
        if not self.ranges or not self._owned(w.data_ptr()):
            return None
        key = (w.data_ptr(), int(w_sn), int(w_sk), int(K), int(N))
        e = self.entries.get(key)
        if e is None:
            e = dict(buf=torch.empty(nbytes // 4, device=w.device, dtype=torch.float32), nt=None, gver=-1, tver=-1, w=w)
            self.entries[key] = e
            self._items_n = -1
        if e["nt"] and e["gver"] == self.gver and e["tver"] == w._version:
            return e["buf"], e["nt"]
        e["gver"], e["tver"], e["w"] = self.gver, w._version, w       # this call packs into the persistent buffer
        return e["buf"], 0

    def set_plan(self, w, w_sn, w_sk, K, N, nt):
        """The channel tile the first call of this (weight, layout) planned (cf_pw_plan_nt): what pack_many packs with."""
        e = self.entries.get((w.data_ptr(), int(w_sn), int(w_sk), int(K), int(N)))
        if e is not None and e["nt"] is None:
            e["nt"] = int(nt)
            self._items_n = -1

    def weights_changed(self):
        """Called by FlatTrainer after the SGD kernel: every registered weight changed; re-pack all known ones in one launch."""
        self.gver += 1
        if not self.entries:
            return
        live = [(k, e) for k, e in self.entries.items() if e["nt"]]      # (nt == 0: the call does not take the persistent kernel)
        if not live:
            return
        if self._items_n != len(live):
            Item = STRUCTS["cf_pack_item"]
            arr = (Item * len(live))()
            for i, ((p_, sn, sk, K, N), e) in enumerate(live):
                arr[i].w, arr[i].pack, arr[i].w_sn, arr[i].w_sk = p_, e["buf"].data_ptr(), sn, sk
                arr[i].K, arr[i].N, arr[i].nt, arr[i].pad = K, N, e["nt"], 0
            raw = bytes(arr)
            dev = next(iter(self.entries.values()))["buf"].device
            self._items = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            self._items_n = len(live)
        call("cf_pw_pack_many", ptr(self._items), self._items_n, stream_ptr())
        for _, e in live:
            e["gver"], e["tver"] = self.gver, e["w"]._version


PACKS = PackCache()


def pw_conv(x, w, y, B, K, N, g, *, w_sn=None, w_sk=1, x2=None, bias=None, pro=PRO_NONE, pro_tabs=(None, None, None),
            epi=EPI_NONE, aux=None, epi_tabs=(None, None), stats=None, stats_mode=STATS_NONE, gather_in=0, scatter_out=0,
            accumulate=0, tc=None):
    wpack, wbytes, wnt, hit = None, 0, 0, None
    strided_1x1 = (g.kt * g.kh * g.kw == 1 and g.ch_stride == 1 and g.pt == 0 and g.ph == 0 and g.pw == 0)
    sn = K if w_sn is None else w_sn
    if (USE_TC if tc is None else tc) and ((not gather_in and not scatter_out) or strided_1x1):
        wbytes = int(lib.cf_pw_tc_ws_bytes(K, N))
        hit = PACKS.lookup(w, sn, w_sk, K, N, wbytes)
        if hit is not None:
            wpack, wnt = hit
        else:
            wpack = torch.empty(wbytes // 4, device=y.device, dtype=torch.float32)
    a = make("cf_pw_args", x=x, x2=x2, w=w, bias=bias, y=y, pro_a=pro_tabs[0], pro_b=pro_tabs[1], pro_c=pro_tabs[2],
             aux=aux, epi_a=epi_tabs[0], epi_b=epi_tabs[1], stats=stats, w_sn=sn, w_sk=w_sk,
             B=B, K=K, N=N, g=g, gather_in=gather_in, scatter_out=scatter_out, accumulate=accumulate, pro_mode=pro,
             epi_mode=epi, stats_mode=stats_mode, wpack=wpack, wpack_bytes=wbytes, wpack_nt=wnt)
    if hit is not None and wnt == 0:
        import ctypes
        PACKS.set_plan(w, sn, w_sk, K, N, lib.cf_pw_plan_nt(ctypes.byref(a)))
    call_struct("cf_pw_conv", a)
    return y


def pw_wgrad(dy, x, dw, B, K, N, g, *, dy2=None, dy_mode=PRO_NONE, dy_tabs=(None, None, None), x_mode=PRO_NONE,
             x_tabs=(None, None), dbias=None, gather_in=0):
    a = make("cf_pw_wgrad_args", dy=dy, dy2=dy2, dy_a=dy_tabs[0], dy_b=dy_tabs[1], dy_c=dy_tabs[2], x=x, x_a=x_tabs[0],
             x_b=x_tabs[1], dw=dw, dbias=dbias, B=B, K=K, N=N, g=g, gather_in=gather_in, dy_mode=dy_mode, x_mode=x_mode)
    call_struct("cf_pw_wgrad", a)
    return dw


def dw_call(fn, x, w, y, B, C, g, *, x2=None, pro=PRO_NONE, pro_tabs=(None, None, None), aux=None, epi=EPI_NONE,
            epi_tabs=(None, None), stats=None, stats_mode=STATS_NONE, dw_out=None):
    a = make("cf_dw_args", x=x, x2=x2, w=w, y=y, pro_a=pro_tabs[0], pro_b=pro_tabs[1], pro_c=pro_tabs[2], aux=aux,
             epi_a=epi_tabs[0], epi_b=epi_tabs[1], stats=stats, dw_out=dw_out, B=B, C=C, g=g, pro_mode=pro, epi_mode=epi,
             stats_mode=stats_mode)
    call_struct(fn, a)
    return y


# Side streams for independent small-kernel chains (the five Rewight branches of the fusion block: a few thousand rows each,
# launch-latency-bound one after the other, concurrent when forked).  Weight-gradient kernels write straight into the flat
# gradient buffer, which autograd does not track: whoever consumes the gradients joins the side streams first
# (train.FlatTrainer.allreduce does; bench.py joins before it ends a CUDA-graph capture).
_SIDE = {}


def side_streams(device, n):
    key = str(device)
    pool = _SIDE.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def join_side_streams():
    """Make the current stream wait for everything enqueued on the side streams so far.  Inside a CUDA-graph capture only the
    side streams that were forked into the capture are waited for (waiting for a stream outside the capture would invalidate
    it; such a stream holds no work of the captured step)."""
    if not _SIDE:
        return
    cur = torch.cuda.current_stream()
    capturing = torch.cuda.is_current_stream_capturing()
    for s in _SIDE.get(str(cur.device), []):
        if capturing:
            with torch.cuda.stream(s):
                if not torch.cuda.is_current_stream_capturing():
                    continue
        cur.wait_stream(s)


WGRAD_STREAM_ROWS = 1 << 16    # Bottlenecks with at most this many output rows per launch (B*T*H*W) run their pointwise weight
                               # gradients on a side stream, joined before the block's backward returns; 0 = never.  Larger
                               # layers fill the GPU with either chain and only time-slice (profiles/r02_ab_same_box.md)


class _WgradStream:
    """`with _WgradStream(dev) as ws: ws.run(fn)`: fn's kernels go to a dedicated side stream that first waits for the work
    enqueued on the current stream so far; __exit__ joins.  Everything the side kernels read is kept alive by the caller's
    locals until the join, so no allocator bookkeeping (record_stream) is needed."""

    def __init__(self, device, enabled):
        self.on = bool(enabled) and torch.device(device).type == "cuda"
        if self.on:
            self.main = torch.cuda.current_stream()
            self.side = side_streams(device, 6)[5]

    def __enter__(self):
        return self

    def run(self, fn):
        if not self.on:
            return fn()
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            return fn()

    def __exit__(self, *exc):
        if self.on:
            self.main.wait_stream(self.side)
        return False


class StatsArena:
    """One zero-filled fp64 buffer for all the BatchNorm statistics / backward sums of a forward pass: ONE memset per
    network pass instead of two per block (the fp64 atomics of the producer epilogues accumulate into slices of it).
    A slice handed out for the backward pass is consumed once (no double backward through these Functions)."""

    def __init__(self, n, device):
        self.buf = torch.zeros(max(int(n), 1), device=device, dtype=torch.float64)
        self.off = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= d
        if self.off + n > self.buf.numel():                    # sized from the module tree; never expected
            return torch.zeros(shape, device=self.buf.device, dtype=torch.float64)
        out = self.buf[self.off:self.off + n].view(shape)
        self.off += n
        return out


def _zeros64(cfg, *shape, device=None):
    arena = getattr(cfg, "arena", None)
    if arena is not None:
        return arena.take(*shape)
    return torch.zeros(shape, device=device, dtype=torch.float64)


class BNCfg:
    """What the kernels need to know about one SubBatchNorm3d (x3d_fine.py:13-62)."""

    def __init__(self, module):
        self.m = module

    @property
    def splits(self):
        return self.m.num_splits


def bn_finalize(stats, bn, B, C, rows, training, device):
    """-> (tab_a, tab_b, mean, invstd); updates split_bn running statistics in training.
    Eval-mode tables depend only on the running statistics and the affine parameters: they are cached on the module and
    rebuilt when one of those tensors changes (autograd version counters; aggregate_stats() re-assigns bn.running_*)."""
    m = bn.m
    if not training:
        key = (B, str(device), m.bn.running_mean.data_ptr(), m.bn.running_var.data_ptr(), m.bn.running_mean._version,
               m.bn.running_var._version, m.weight._version, m.bias._version, m.weight.data_ptr())
        hit = getattr(m, "_cf_eval_tabs", None)
        if hit is not None and hit[0] == key and not torch.cuda.is_current_stream_capturing():
            return hit[1]
    splits = m.num_splits if training else 1
    tabs = torch.empty(2, B, C, device=device, dtype=torch.float32)
    ms = torch.empty(2, splits, C, device=device, dtype=torch.float32)
    if training:
        rm, rv = m.split_bn.running_mean, m.split_bn.running_var
        m._pending_batches = getattr(m, "_pending_batches", 0) + 1
    else:
        rm, rv = m.bn.running_mean, m.bn.running_var
    a = make("cf_bn_args", stats=stats, gamma=m.weight, beta=m.bias, running_mean=rm, running_var=rv, tab_a=tabs[0],
             tab_b=tabs[1], mean=ms[0], invstd=ms[1], B=B, C=C, splits=splits, rows_per_sample=rows,
             momentum=float(m.split_bn.momentum if m.split_bn.momentum is not None else 0.1), eps=float(m.split_bn.eps),
             training=int(training))
    call_struct("cf_bn_finalize", a)
    if not training:
        m._cf_eval_tabs = (key, (tabs[0], tabs[1], ms[0], ms[1]))
    return tabs[0], tabs[1], ms[0], ms[1]


def bn_bwd_coeffs(sums, gamma, mean, invstd, dgamma, dbeta, B, C, rows, training, gate=None, cst=None):
    tabs = torch.empty(3, B, C, device=sums.device, dtype=torch.float32)
    a = make("cf_bn_bwd_args", sums=sums, gamma=gamma, mean=mean, invstd=invstd, gate=gate, cst=cst, dgamma=dgamma,
             dbeta=dbeta, tab_p=tabs[0], tab_q=tabs[1], tab_r=tabs[2], B=B, C=C, splits=mean.shape[0],
             rows_per_sample=rows, training=int(training))
    call_struct("cf_bn_bwd_coeffs", a)
    return tabs[0], tabs[1], tabs[2]


def residual_fwd(y, ta, tb, out, B, C, rows, res=None, ra=None, rb=None, pooled=None, pool_geom=(0, 0, 0, 0, 0)):
    T, H, W, rh, rw = pool_geom
    a = make("cf_residual_args", y=y, tab_a=ta, tab_b=tb, res=res, res_a=ra, res_b=rb, out=out, pooled=pooled, B=B, C=C,
             T=T, H=H, W=W, rh=rh, rw=rw, rows_per_sample=rows)
    call_struct("cf_residual_fwd", a)
    return out


def residual_bwd(dout, out, y, dz, sums_y, B, C, rows, res=None, sums_res=None, dpool=None, pool_geom=(0, 0, 0, 0, 0)):
    T, H, W, rh, rw = pool_geom
    a = make("cf_residual_bwd_args", dout=dout, out=out, y=y, res=res, dpool=dpool, dz=dz, sums_y=sums_y, sums_res=sums_res,
             B=B, C=C, T=T, H=H, W=W, rh=rh, rw=rw, rows_per_sample=rows)
    call_struct("cf_residual_bwd", a)
    return dz


def _flat_grads(params, device):
    """Gradient targets for a list of parameters (None entries allowed) -> (buffers, returns).

    ``buffers[i]`` is what the weight-gradient kernels accumulate (+=) into; ``returns[i]`` is what
    the autograd Function hands back.  A parameter owned by a ``train.FlatTrainer`` carries
    ``_cf_grad``, a view of the flat gradient buffer (zeroed by the fused SGD kernel): kernels
    accumulate into it directly and autograd gets None -- no per-parameter AccumulateGrad kernel, no
    memset.  Any other parameter gets a view of one zero-filled scratch buffer (a single memset)."""
    own = [p for p in params if p is not None and getattr(p, "_cf_grad", None) is None]
    flat = torch.zeros(sum(p.numel() for p in own), device=device, dtype=torch.float32) if own else None
    bufs, rets, o = [], [], 0
    for p in params:
        if p is None:
            bufs.append(None)
            rets.append(None)
        elif getattr(p, "_cf_grad", None) is not None:
            bufs.append(p._cf_grad)
            rets.append(None)
        else:
            n = p.numel()
            v = flat[o:o + n].view(p.shape)
            o += n
            bufs.append(v)
            rets.append(v)
    return bufs, rets


# ----------------------------------------------------------------------------------------
class BottleneckFn(torch.autograd.Function):
    """Bottleneck.forward (x3d_fine.py:146-175): conv1->bn1->relu->conv2(dw)->bn2->[SE]->swish->
    conv3->bn3->(+res)->relu, as 7-10 kernel launches forward and 12-16 backward."""

    @staticmethod
    def forward(ctx, x, cfg, w1, g1, b1, w2, g2, b2, w3, g3, b3, fw1, fb1, fw2, fb2, wd, gd, bd):
        x = cl(x)
        dev = x.device
        B, Cin, T, H, W = x.shape
        Ce, Co = w1.shape[0], w3.shape[0]
        s, ts = cfg.stride, cfg.t_stride
        To, Ho, Wo = (T - 1) // ts + 1, (H - 1) // s + 1, (W - 1) // s + 1
        Rin, Rout = T * H * W, To * Ho * Wo
        tr = cfg.training
        has_se, has_ds = fw1 is not None, wd is not None
        Cmax = max(Ce, Co)
        if not tr and getattr(cfg, "inference", False):
            return BottleneckFn._inference(x, cfg, w1, w2, w3, fw1, fb1, fw2, fb2, wd, (B, Cin, Ce, Co, T, H, W, To, Ho, Wo, s, ts))
        need_stats = tr or has_se                      # eval: only the SE pool needs sum(y2)
        stats = _zeros64(cfg, 4, B, Cmax, 2, device=dev) if need_stats else None
        st = (lambda i, C: stats[i].view(-1)[: B * C * 2]) if need_stats else (lambda i, C: None)
        ctx.sums = _zeros64(cfg, 4, B, Cmax, 2, device=dev) if getattr(cfg, "arena", None) is not None else None
        smode = STATS_SUM_SQ if tr else STATS_NONE
        g_in, g_out = geom(T, H, W), geom(To, Ho, Wo)
        g_dw = geom(To, Ho, Wo, T, H, W, k=(3, 3, 3), s=(ts, s, s), p=(1, 1, 1))
        g_ds = geom(To, Ho, Wo, T, H, W, s=(ts, s, s), pos_stride=Cin, ch_stride=1, sample_stride=Rin * Cin)

        y1 = new_act(B, Ce, T, H, W, dev)
        pw_conv(x, w1, y1, B, Cin, Ce, g_in, stats=st(0, Ce) if tr else None, stats_mode=smode)
        a1, bb1, m1, i1 = bn_finalize(st(0, Ce) if tr else None, cfg.bn1, B, Ce, Rin, tr, dev)
        y2 = new_act(B, Ce, To, Ho, Wo, dev)
        dw_call("cf_dw_conv_fwd", y1, w2, y2, B, Ce, g_dw, pro=PRO_AFFINE_RELU, pro_tabs=(a1, bb1, None),
                stats=st(1, Ce), stats_mode=STATS_SUM_SQ if need_stats else STATS_NONE)
        a2, bb2, m2, i2 = bn_finalize(st(1, Ce) if tr else None, cfg.bn2, B, Ce, Rout, tr, dev)
        se_saved = None
        ga, gb = a2, bb2
        if has_se:
            Wd = fw1.shape[0]
            sv = torch.empty(4, B, Ce, device=dev, dtype=torch.float32)     # pooled, gate, out_a, out_b
            hid = torch.empty(B, Wd, device=dev, dtype=torch.float32)
            a = make("cf_se_args", stats=st(1, Ce), tab_a=a2, tab_b=bb2, w1=fw1, b1=fb1, w2=fw2, b2=fb2, pooled=sv[0],
                     hidden=hid, gate=sv[1], out_a=sv[2], out_b=sv[3], B=B, C=Ce, Wd=Wd, rows_per_sample=Rout)
            call_struct("cf_se_fwd", a)
            ga, gb = sv[2], sv[3]
            se_saved = (sv[0], hid, sv[1])
        y3 = new_act(B, Co, To, Ho, Wo, dev)
        pw_conv(y2, w3, y3, B, Ce, Co, g_out, pro=PRO_AFFINE_SWISH, pro_tabs=(ga, gb, None), stats=st(2, Co) if tr else None,
                stats_mode=smode)
        a3, bb3, m3, i3 = bn_finalize(st(2, Co) if tr else None, cfg.bn3, B, Co, Rout, tr, dev)
        yd = ad = bbd = md = idd = None
        if has_ds:
            yd = new_act(B, Co, To, Ho, Wo, dev)
            pw_conv(x, wd, yd, B, Cin, Co, g_ds, gather_in=1, stats=st(3, Co) if tr else None, stats_mode=smode)
            ad, bbd, md, idd = bn_finalize(st(3, Co) if tr else None, cfg.bnd, B, Co, Rout, tr, dev)
        out = new_act(B, Co, To, Ho, Wo, dev)
        pool = getattr(cfg, "pool", None)                     # (rh, rw): stage-final block of the global tower
        pooled = None
        if pool is not None:
            pooled = new_act(B, Co, To, Ho // pool[0], Wo // pool[1], dev)
            ctx.pool_geom = (To, Ho, Wo, pool[0], pool[1])
        residual_fwd(y3, a3, bb3, out, B, Co, Rout, res=yd if has_ds else x, ra=ad, rb=bbd, pooled=pooled,
                     pool_geom=ctx.pool_geom if pool is not None else (0, 0, 0, 0, 0))
        ctx.has_pool = pool is not None

        ctx.cfg = cfg
        ctx.dims = (B, Cin, Ce, Co, T, H, W, To, Ho, Wo, s, ts)
        ctx.se_saved = se_saved
        ctx.aux = (a1, bb1, m1, i1, a2, bb2, m2, i2, ga, gb, a3, bb3, m3, i3, ad, bbd, md, idd, st(1, Ce))
        ctx.save_for_backward(x, y1, y2, y3, yd, out, w1, g1, w2, g2, w3, g3, fw1, fw2, wd, gd)
        ctx.param_shapes = [p for p in (w1, g1, b1, w2, g2, b2, w3, g3, b3, fw1, fb1, fw2, fb2, wd, gd, bd)]
        if pool is not None:
            return out, pooled
        return out

    @staticmethod
    def _inference(x, cfg, w1, w2, w3, fw1, fb1, fw2, fb2, wd, dims):
        """Eval mode without autograd (validation, feature extraction): every BatchNorm is a constant affine map, so the
        block is THREE conv launches -- conv1 | depthwise (bn1+ReLU prologue) | conv3 (bn2 [+SE] + Swish prologue, and bn3 +
        shortcut + ReLU folded into its epilogue: x3d_fine.py:167-173) -- plus the SE gate and, in the first block of a stage,
        the shortcut conv with its BatchNorm folded into ITS epilogue.  The training-mode sequence needs y3 and a separate
        residual join because bn3's batch statistics only exist after conv3 has finished; here nothing is materialised
        between conv3 and the block output, and nothing is saved."""
        B, Cin, Ce, Co, T, H, W, To, Ho, Wo, s, ts = dims
        dev = x.device
        has_se, has_ds = fw1 is not None, wd is not None
        Rin, Rout = T * H * W, To * Ho * Wo
        g_in, g_out = geom(T, H, W), geom(To, Ho, Wo)
        g_dw = geom(To, Ho, Wo, T, H, W, k=(3, 3, 3), s=(ts, s, s), p=(1, 1, 1))
        y1 = new_act(B, Ce, T, H, W, dev)
        pw_conv(x, w1, y1, B, Cin, Ce, g_in)
        a1, bb1, _, _ = bn_finalize(None, cfg.bn1, B, Ce, Rin, False, dev)
        y2 = new_act(B, Ce, To, Ho, Wo, dev)
        stats = _zeros64(cfg, B, Ce, 2, device=dev) if has_se else None          # SE pooling: sum of the raw conv2 output
        dw_call("cf_dw_conv_fwd", y1, w2, y2, B, Ce, g_dw, pro=PRO_AFFINE_RELU, pro_tabs=(a1, bb1, None), stats=stats,
                stats_mode=STATS_SUM_SQ if has_se else STATS_NONE)
        ga, gb, _, _ = bn_finalize(None, cfg.bn2, B, Ce, Rout, False, dev)
        if has_se:
            Wd = fw1.shape[0]
            sv = torch.empty(4, B, Ce, device=dev, dtype=torch.float32)
            hid = torch.empty(B, Wd, device=dev, dtype=torch.float32)
            call_struct("cf_se_fwd", make("cf_se_args", stats=stats.view(-1), tab_a=ga, tab_b=gb, w1=fw1, b1=fb1, w2=fw2, b2=fb2,
                                          pooled=sv[0], hidden=hid, gate=sv[1], out_a=sv[2], out_b=sv[3], B=B, C=Ce, Wd=Wd,
                                          rows_per_sample=Rout))
            ga, gb = sv[2], sv[3]
        shortcut = x
        if has_ds:
            g_ds = geom(To, Ho, Wo, T, H, W, s=(ts, s, s), pos_stride=Cin, ch_stride=1, sample_stride=Rin * Cin)
            ad, bbd, _, _ = bn_finalize(None, cfg.bnd, B, Co, Rout, False, dev)
            shortcut = new_act(B, Co, To, Ho, Wo, dev)
            pw_conv(x, wd, shortcut, B, Cin, Co, g_ds, gather_in=1, epi=EPI_AFFINE, epi_tabs=(ad, bbd))
        a3, bb3, _, _ = bn_finalize(None, cfg.bn3, B, Co, Rout, False, dev)
        out = new_act(B, Co, To, Ho, Wo, dev)
        pw_conv(y2, w3, out, B, Ce, Co, g_out, pro=PRO_AFFINE_SWISH, pro_tabs=(ga, gb, None), epi=EPI_AFFINE_ADD_RELU, aux=shortcut,
                epi_tabs=(a3, bb3))
        pool = getattr(cfg, "pool", None)
        if pool is not None:                                       # stage-final block of the global tower (x3d_fine.py:345-354)
            pooled = new_act(B, Co, To, Ho // pool[0], Wo // pool[1], dev)
            call_struct("cf_block_avgpool_fwd", make("cf_pool_args", x=out, y=pooled, tab_a=None, tab_b=None, B=B, C=Co, T=To, H=Ho,
                                                     W=Wo, rh=pool[0], rw=pool[1]))
            return out, pooled
        return out

    @staticmethod
    def backward(ctx, dout, dpooled=None):
        x, y1, y2, y3, yd, out, w1, g1, w2, g2, w3, g3, fw1, fw2, wd, gd = ctx.saved_tensors
        (a1, bb1, m1, i1, a2, bb2, m2, i2, ga, gb, a3, bb3, m3, i3, ad, bbd, md, idd, stats2) = ctx.aux
        B, Cin, Ce, Co, T, H, W, To, Ho, Wo, s, ts = ctx.dims
        cfg = ctx.cfg
        tr = cfg.training
        dev = x.device
        has_se, has_ds = fw1 is not None, wd is not None
        Rin, Rout = T * H * W, To * Ho * Wo
        dout = cl(dout) if dout is not None else None
        dpooled = cl(dpooled) if (ctx.has_pool and dpooled is not None) else None
        grads, rets = _flat_grads(ctx.param_shapes, dev)
        (dw1, dg1, db1, dw2, dg2, db2, dw3, dg3, db3, dfw1, dfb1, dfw2, dfb2, dwd, dgd, dbd) = grads
        Cmax = max(Ce, Co)
        sums = ctx.sums if ctx.sums is not None else torch.zeros(4, B, Cmax, 2, device=dev, dtype=torch.float64)
        ctx.sums = None
        sm = lambda i, C: sums[i].view(-1)[: B * C * 2]
        g_in, g_out = geom(T, H, W), geom(To, Ho, Wo)
        g_dw = geom(To, Ho, Wo, T, H, W, k=(3, 3, 3), s=(ts, s, s), p=(1, 1, 1))
        g_ds = geom(To, Ho, Wo, T, H, W, s=(ts, s, s), pos_stride=Cin, ch_stride=1, sample_stride=Rin * Cin)

        # join: dz3 = dout*[out>0]; sums vs y3 (bn3) and vs yd (downsample bn)
        dz3 = torch.empty_like(out)
        residual_bwd(dout, out, y3, dz3, sm(0, Co), B, Co, Rout, res=yd, sums_res=sm(3, Co) if has_ds else None, dpool=dpooled,
                     pool_geom=ctx.pool_geom if dpooled is not None else (0, 0, 0, 0, 0))
        P3, Q3, R3 = bn_bwd_coeffs(sm(0, Co), g3, m3, i3, dg3, db3, B, Co, Rout, tr)
        # The weight gradients of the three pointwise convs are leaves of the step (nothing reads them before the optimizer):
        # they run on a side stream next to the data-gradient chain and are joined before this block's backward returns.
        with _WgradStream(dev, B * Rout <= WGRAD_STREAM_ROWS) as ws:
            # conv3
            ws.run(lambda: pw_wgrad(dz3, y2, dw3, B, Ce, Co, g_out, dy2=y3, dy_mode=PRO_AFFINE2, dy_tabs=(P3, Q3, R3),
                                    x_mode=PRO_AFFINE_SWISH, x_tabs=(ga, gb)))
            dU = torch.empty_like(y2)
            pw_conv(dz3, w3, dU, B, Co, Ce, g_out, w_sn=1, w_sk=Ce, x2=y3, pro=PRO_AFFINE2, pro_tabs=(P3, Q3, R3), epi=EPI_DSWISH,
                    aux=y2, epi_tabs=(ga, gb), stats=sm(1, Ce), stats_mode=STATS_SUM_AUX)
            gate = cst = None
            if has_se:
                pooled, hid, gate = ctx.se_saved
                cst = torch.empty(B, Ce, device=dev, dtype=torch.float32)
                a = make("cf_se_bwd_args", sums=sm(1, Ce), stats_y=stats2, tab_a=a2, tab_b=bb2, w1=fw1, w2=fw2, pooled=pooled,
                         hidden=hid, gate=gate, dw1=dfw1, db1=dfb1, dw2=dfw2, db2=dfb2, cst=cst, B=B, C=Ce, Wd=fw1.shape[0],
                         rows_per_sample=Rout)
                call_struct("cf_se_bwd", a)
            P2, Q2, R2 = bn_bwd_coeffs(sm(1, Ce), g2, m2, i2, dg2, db2, B, Ce, Rout, tr, gate=gate, cst=cst)
            # conv2 (depthwise): data gradient and weight gradient from one call (one pass over dU, y2, y1 for the stride-1 convs)
            dz1 = torch.empty_like(y1)
            dw_call("cf_dw_conv_dgrad", dU, w2, dz1, B, Ce, g_dw, x2=y2, pro=PRO_AFFINE2, pro_tabs=(P2, Q2, R2), aux=y1,
                    epi=EPI_DRELU, epi_tabs=(a1, bb1), stats=sm(2, Ce), stats_mode=STATS_SUM_AUX, dw_out=dw2)
            P1, Q1, R1 = bn_bwd_coeffs(sm(2, Ce), g1, m1, i1, dg1, db1, B, Ce, Rin, tr)
            # conv1
            ws.run(lambda: pw_wgrad(dz1, x, dw1, B, Cin, Ce, g_in, dy2=y1, dy_mode=PRO_AFFINE2, dy_tabs=(P1, Q1, R1)))
            dx = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                pw_conv(dz1, w1, dx, B, Ce, Cin, g_in, w_sn=1, w_sk=Cin, x2=y1, pro=PRO_AFFINE2, pro_tabs=(P1, Q1, R1),
                        epi=EPI_NONE if has_ds else EPI_ADD_AUX, aux=None if has_ds else dz3)
            if has_ds:
                Pd, Qd, Rd = bn_bwd_coeffs(sm(3, Co), gd, md, idd, dgd, dbd, B, Co, Rout, tr)
                ws.run(lambda: pw_wgrad(dz3, x, dwd, B, Cin, Co, g_ds, dy2=yd, dy_mode=PRO_AFFINE2, dy_tabs=(Pd, Qd, Rd), gather_in=1))
                if dx is not None:
                    pw_conv(dz3, wd, dx, B, Co, Cin, g_ds, w_sn=1, w_sk=Cin, x2=yd, pro=PRO_AFFINE2, pro_tabs=(Pd, Qd, Rd),
                            scatter_out=1, accumulate=1)
        return (dx, None, *rets)


# ----------------------------------------------------------------------------------------
class StemFn(torch.autograd.Function):
    """conv1_s (1x3x3, stride (1,2,2), 3->24) -> conv1_t (5x1x1 depthwise) -> bn1 -> relu
    (x3d_fine.py:210-223, 334-337).  Input is the network's NCTHW clip, output channels-last."""

    @staticmethod
    def forward(ctx, x, cfg, ws, wt, gamma, beta):
        x = x.float()
        dev = x.device
        B, Ci, T, H, W = x.shape
        if not (x.stride(4) == 1 and x.stride(3) == W and x.stride(2) == H * W):
            x = x.contiguous()           # a temporal window x_full[:, :, a:b] of an NCTHW clip is gathered in place
        C = ws.shape[0]
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        R = T * Ho * Wo
        tr = cfg.training
        g_s = geom(T, Ho, Wo, T, H, W, k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1), pos_stride=1, ch_stride=x.stride(1),
                   sample_stride=x.stride(0))
        g_t = geom(T, Ho, Wo, k=(5, 1, 1), p=(2, 0, 0))
        y0 = new_act(B, C, T, Ho, Wo, dev)
        pw_conv(x, ws, y0, B, Ci * 9, C, g_s, gather_in=1)
        yt = new_act(B, C, T, Ho, Wo, dev)
        stats = _zeros64(cfg, B, C, 2, device=dev) if tr else None
        ctx.sums = _zeros64(cfg, B, C, 2, device=dev) if getattr(cfg, "arena", None) is not None else None
        dw_call("cf_dw_conv_fwd", y0, wt, yt, B, C, g_t, stats=stats, stats_mode=STATS_SUM_SQ if tr else STATS_NONE)
        a, b, m, i = bn_finalize(stats, cfg.bn1, B, C, R, tr, dev)
        out = new_act(B, C, T, Ho, Wo, dev)
        residual_fwd(yt, a, b, out, B, C, R)
        ctx.cfg, ctx.dims, ctx.geoms, ctx.bn = cfg, (B, Ci, C, T, H, W, Ho, Wo), (g_s, g_t), (m, i)
        ctx.save_for_backward(x, y0, yt, out, ws, wt, gamma)
        ctx.params = (ws, wt, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y0, yt, out, ws, wt, gamma = ctx.saved_tensors
        B, Ci, C, T, H, W, Ho, Wo = ctx.dims
        g_s, g_t = ctx.geoms
        m, i = ctx.bn
        dev = x.device
        R = T * Ho * Wo
        dout = cl(dout)
        (dws, dwt, dgam, dbet), rets = _flat_grads(ctx.params, dev)
        sums = ctx.sums if ctx.sums is not None else torch.zeros(B, C, 2, device=dev, dtype=torch.float64)
        ctx.sums = None
        dz = torch.empty_like(out)
        residual_bwd(dout, out, yt, dz, sums, B, C, R)
        P, Q, Rr = bn_bwd_coeffs(sums, gamma, m, i, dgam, dbet, B, C, R, ctx.cfg.training)
        dy0 = torch.empty_like(y0)                # conv1_t: data gradient + weight gradient in one march over (dz, yt, y0)
        dw_call("cf_dw_conv_dgrad", dz, wt, dy0, B, C, g_t, x2=yt, pro=PRO_AFFINE2, pro_tabs=(P, Q, Rr), aux=y0, dw_out=dwt)
        pw_wgrad(dy0, x, dws, B, Ci * 9, C, g_s, gather_in=1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.zeros(x.shape, device=dev, dtype=torch.float32)
            g_dx = geom(T, Ho, Wo, T, H, W, k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1), pos_stride=1, ch_stride=T * H * W,
                        sample_stride=Ci * T * H * W)
            pw_conv(dy0, ws, dx, B, C, Ci * 9, g_dx, w_sn=1, w_sk=Ci * 9, scatter_out=1)
        return (dx, None, *rets)


class ConvBNReluPoolFn(torch.autograd.Function):
    """conv5 (1x1x1) -> bn5 -> relu -> block average pool over (H,W)
    (x3d_fine.py:356-366; pool to 7x7 for the global tower, :360)."""

    @staticmethod
    def forward(ctx, x, cfg, w, gamma, beta, rh, rw):
        x = cl(x)
        dev = x.device
        B, Cin, T, H, W = x.shape
        C = w.shape[0]
        R = T * H * W
        tr = cfg.training
        y = new_act(B, C, T, H, W, dev)
        stats = _zeros64(cfg, B, C, 2, device=dev) if tr else None
        ctx.sums = _zeros64(cfg, B, C, 2, device=dev) if getattr(cfg, "arena", None) is not None else None
        pw_conv(x, w, y, B, Cin, C, geom(T, H, W), stats=stats, stats_mode=STATS_SUM_SQ if tr else STATS_NONE)
        a, b, m, i = bn_finalize(stats, cfg.bn, B, C, R, tr, dev)
        out = new_act(B, C, T, H // rh, W // rw, dev)
        call_struct("cf_block_avgpool_fwd", make("cf_pool_args", x=y, y=out, tab_a=a, tab_b=b, B=B, C=C, T=T, H=H, W=W, rh=rh, rw=rw))
        ctx.cfg, ctx.dims, ctx.tabs = cfg, (B, Cin, C, T, H, W, rh, rw), (a, b, m, i)
        ctx.save_for_backward(x, y, w, gamma)
        ctx.params = (w, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, w, gamma = ctx.saved_tensors
        B, Cin, C, T, H, W, rh, rw = ctx.dims
        a, b, m, i = ctx.tabs
        dev = x.device
        R = T * H * W
        dout = cl(dout)
        (dw, dgam, dbet), rets = _flat_grads(ctx.params, dev)
        sums = ctx.sums if ctx.sums is not None else torch.zeros(B, C, 2, device=dev, dtype=torch.float64)
        ctx.sums = None
        dz = torch.empty_like(y)
        call_struct("cf_block_avgpool_bwd", make("cf_pool_bwd_args", dy=dout, x=y, tab_a=a, tab_b=b, dz=dz, sums=sums, B=B, C=C,
                                                 T=T, H=H, W=W, rh=rh, rw=rw, accumulate=0))
        P, Q, Rr = bn_bwd_coeffs(sums, gamma, m, i, dgam, dbet, B, C, R, ctx.cfg.training)
        g = geom(T, H, W)
        pw_wgrad(dz, x, dw, B, Cin, C, g, dy2=y, dy_mode=PRO_AFFINE2, dy_tabs=(P, Q, Rr))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            pw_conv(dz, w, dx, B, C, Cin, g, w_sn=1, w_sk=Cin, x2=y, pro=PRO_AFFINE2, pro_tabs=(P, Q, Rr))
        return (dx, None, *rets, None, None)


class AvgPoolFn(torch.autograd.Function):
    """F.adaptive_avg_pool3d(x, (None, H/rh, W/rw)) for H,W multiples of the output (x3d_fine.py:345-354)."""

    @staticmethod
    def forward(ctx, x, rh, rw):
        x = cl(x)
        B, C, T, H, W = x.shape
        out = new_act(B, C, T, H // rh, W // rw, x.device)
        call_struct("cf_block_avgpool_fwd", make("cf_pool_args", x=x, y=out, tab_a=None, tab_b=None, B=B, C=C, T=T, H=H, W=W, rh=rh, rw=rw))
        ctx.dims = (B, C, T, H, W, rh, rw)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C, T, H, W, rh, rw = ctx.dims
        dout = cl(dout)
        dz = new_act(B, C, T, H, W, dout.device)
        call_struct("cf_block_avgpool_bwd", make("cf_pool_bwd_args", dy=dout, x=None, tab_a=None, tab_b=None, dz=dz, sums=None, B=B,
                                                 C=C, T=T, H=H, W=W, rh=rh, rw=rw, accumulate=0))
        return dz, None, None


ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


class LinearRowsFn(torch.autograd.Function):
    """y[b,r,:] = act(W x[b,r,:] + bias) on a [B,R,K] row tensor: fc1 (1x1x1 conv, no bias, + ReLU)
    and fc2 (nn.Linear) of the head (x3d_fine.py:370-380), and the k=1 Conv1d layers of the fusion
    block (x3d_coarse.py:216-219,232-246,335-336).  act: ACT_NONE / ACT_RELU / ACT_SIGMOID."""

    @staticmethod
    def forward(ctx, x, w, bias, act):
        act = int(act)
        x = x.contiguous().float()
        B, R, K = x.shape
        w2 = w.reshape(w.shape[0], -1)
        N = w2.shape[0]
        y = torch.empty(B, R, N, device=x.device, dtype=torch.float32)
        pw_conv(x, w2, y, B, K, N, geom(R, 1, 1), bias=bias, epi=(EPI_NONE, EPI_RELU, EPI_SIGMOID)[act])
        ctx.act, ctx.dims, ctx.wshape, ctx.has_bias = act, (B, R, K, N), w.shape, bias is not None
        ctx.save_for_backward(x, w2, y if act else None)
        ctx.params = (w, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w2, y = ctx.saved_tensors
        B, R, K, N = ctx.dims
        dy = dy.contiguous().float()
        g = geom(R, 1, 1)
        if ctx.act:                       # dy *= act'(y)
            dya = torch.empty_like(dy)
            call("cf_relu_bwd" if ctx.act == ACT_RELU else "cf_sigmoid_bwd", ptr(dy), ptr(y), ptr(dya), dy.numel(),
                 stream_ptr())
            dy = dya
        (dw, db), rets = _flat_grads(ctx.params, dy.device)
        pw_wgrad(dy, x, dw, B, K, N, g, dbias=db)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            pw_conv(dy, w2, dx, B, N, K, g, w_sn=1, w_sk=K)
        return (dx, *rets, None)


class SwishFn(torch.autograd.Function):
    """SwishEfficient (x3d_fine.py:74-86): saves only x."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous() if not x.is_contiguous(memory_format=CL3) else x
        out = torch.empty_like(x)
        call("cf_swish_fwd", ptr(x), ptr(out), x.numel(), stream_ptr())
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous(memory_format=CL3) if (x.dim() == 5 and not x.is_contiguous()) else dy.contiguous()
        dx = torch.empty_like(x)
        call("cf_swish_bwd", ptr(x), ptr(dy), ptr(dx), x.numel(), stream_ptr())
        return dx


class StandaloneBNFn(torch.autograd.Function):
    """SubBatchNorm3d.forward called on its own (x3d_fine.py:51-62): statistics kernel +
    table + affine apply; backward through the same affine-map formulation."""

    @staticmethod
    def forward(ctx, x, bn, training, gamma, beta):
        x = cl(x)
        B, C, T, H, W = x.shape
        R = T * H * W
        stats = None
        if training:
            stats = torch.zeros(B, C, 2, device=x.device, dtype=torch.float64)
            call("cf_channel_stats", ptr(x), None, ptr(stats), B, C, R, stream_ptr())
        a, b, m, i = bn_finalize(stats, bn, B, C, R, training, x.device)
        out = torch.empty_like(x)
        call_struct("cf_affine_apply", make("cf_affine_args", x=x, x2=None, tab_a=a, tab_b=b, tab_c=None, out=out, B=B, C=C,
                                            rows_per_sample=R, mode=PRO_AFFINE))
        ctx.save_for_backward(x, gamma)
        ctx.params = (gamma, beta)
        ctx.misc = (m, i, training, B, C, R)
        return out

    @staticmethod
    def backward(ctx, dz):
        x, gamma = ctx.saved_tensors
        m, i, training, B, C, R = ctx.misc
        dz = cl(dz)
        sums = torch.zeros(B, C, 2, device=x.device, dtype=torch.float64)
        call("cf_channel_stats", ptr(dz), ptr(x), ptr(sums), B, C, R, stream_ptr())
        (dgam, dbet), rets = _flat_grads(ctx.params, x.device)
        P, Q, Rr = bn_bwd_coeffs(sums, gamma, m, i, dgam, dbet, B, C, R, training)
        dx = torch.empty_like(x)
        call_struct("cf_affine_apply", make("cf_affine_args", x=dz, x2=x, tab_a=P, tab_b=Q, tab_c=Rr, out=dx, B=B, C=C,
                                            rows_per_sample=R, mode=PRO_AFFINE2))
        return (dx, None, None, *rets)
