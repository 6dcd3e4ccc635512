"""Batched 1-D linear interpolation -- drop-in for the reference's ``interp1d.py``.

``Interp1d()(x, y, xnew, out=None)`` has the reference's calling convention and shape rules
(interp1d.py:5-60): inputs are 1-D or 2-D, a single row broadcasts over the rows of the others.
Like the reference (whose ``__call__`` invokes ``forward`` directly, so autograd differentiates
the interpolation formula), gradients flow to x, y and xnew.  CUDA only: cf_interp1d_fwd/bwd."""
import torch

from ._lib import call, ptr, stream_ptr


class _Interp1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, xnew):
        D = max(x.shape[0], y.shape[0], xnew.shape[0])
        N, P = x.shape[1], xnew.shape[1]
        rs = lambda t: 0 if (t.shape[0] == 1 and D > 1) else t.shape[1]
        ynew = torch.empty(D, P, device=x.device, dtype=torch.float32)
        ind = torch.empty(D, P, device=x.device, dtype=torch.int32)
        call("cf_interp1d_fwd", ptr(x), ptr(y), ptr(xnew), ptr(ynew), ptr(ind), D, N, P, rs(x), rs(y), rs(xnew), stream_ptr())
        ctx.save_for_backward(x, y, xnew, ind)
        ctx.dims = (D, N, P, rs(x), rs(y), rs(xnew))
        ctx.mark_non_differentiable(ind)
        return ynew, ind

    @staticmethod
    def backward(ctx, dynew, _dind):
        x, y, xnew, ind = ctx.saved_tensors
        D, N, P, xrs, yrs, qrs = ctx.dims
        need = ctx.needs_input_grad
        dx = torch.zeros_like(x) if need[0] else None
        dy = torch.zeros_like(y) if need[1] else None
        dq = torch.zeros_like(xnew) if need[2] else None
        call("cf_interp1d_bwd", ptr(x), ptr(y), ptr(xnew), ptr(ind), ptr(dynew.contiguous()), ptr(dx), ptr(dy), ptr(dq), D, N,
             P, xrs, yrs, qrs, stream_ptr())
        return dx, dy, dq


class Interp1d:
    """Same call surface as the reference class (interp1d.py:4-6): ``Interp1d()(x, y, xnew, out)``."""

    def __call__(self, x, y, xnew, out=None):
        return self.forward(x, y, xnew, out)

    @staticmethod
    def forward(x, y, xnew, out=None):
        v = {}
        for name, vec in (("x", x), ("y", y), ("xnew", xnew)):
            assert vec.dim() <= 2, "interp1d: all inputs must be at most 2-D."
            v[name] = (vec[None, :] if vec.dim() == 1 else vec).contiguous().float()
        assert len({str(t.device) for t in v.values()}) == 1, "All parameters must be on the same device."
        assert (v["x"].shape[1] == v["y"].shape[1]
                and (v["x"].shape[0] == v["y"].shape[0] or v["x"].shape[0] == 1 or v["y"].shape[0] == 1)), (
            "x and y must have the same number of columns, and either the same number of row or one of them having "
            "only one row.")
        if v["x"].shape[0] == 1 and v["y"].shape[0] > 1:
            raise NotImplementedError("a single x row with several y rows (flat slope indexing, interp1d.py:130) is not built")
        shape = None
        if v["x"].shape[0] == 1 and v["y"].shape[0] == 1 and v["xnew"].shape[0] > 1:
            shape = v["xnew"].shape                     # one problem for all query rows (interp1d.py:62-70)
            v["xnew"] = v["xnew"].reshape(1, -1)
        ynew, _ = _Interp1dFn.apply(v["x"], v["y"], v["xnew"])
        # `out` is accepted for signature parity; the reference rebinds its result and never fills it
        return ynew.view(shape) if shape is not None else ynew
