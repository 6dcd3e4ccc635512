"""APMeter with the reference's surface (apmeter.py:7-136: reset / add(output, target, weight=None) / value())
for the evaluation loops of train_coarse_fineFEAT.py:241-263 and extract_fineFEAT.py.

Scores and targets stay on the GPU (the reference moves every batch to numpy on the host,
train_coarse_fineFEAT.py:241-242,262-263); value() sorts all classes at once (class-major [K,N], descending, stable)
and runs the scan / divide / masked-sum kernel cf_ap_sorted -- one CTA per class."""
import numpy as np
import torch

from ._lib import call, ptr, stream_ptr


class APMeter:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.reset()

    def reset(self):
        self._scores, self._targets, self._weights = [], [], []

    def add(self, output, target, weight=None):
        """output [N,K] scores, target [N,K] binary, weight [N] (> 0) or None; tensors or numpy arrays (apmeter.py:31-95)."""
        as_t = lambda v: torch.from_numpy(np.ascontiguousarray(v)) if not torch.is_tensor(v) else v
        output, target = as_t(output), as_t(target)
        if output.dim() == 1:
            output = output.view(-1, 1)
        if target.dim() == 1:
            target = target.view(-1, 1)
        assert output.dim() == 2 and target.dim() == 2, "wrong output / target size (should be 1D or 2D with one column per class)"
        assert output.shape == target.shape
        if self._scores:
            assert target.shape[1] == self._targets[0].shape[1], "dimensions for output should match previously added examples."
        target = target.to(self.device, torch.float32)
        assert bool(((target == 0) | (target == 1)).all()), "targets should be binary (0 or 1)"
        if weight is not None:
            weight = as_t(weight).squeeze()
            assert weight.dim() == 1 and weight.numel() == target.shape[0], "Weight dimension 1 should be the same as that of target"
            assert float(weight.min()) >= 0, "Weight should be non-negative only"
            self._weights.append(weight.to(self.device, torch.float32))
        assert (weight is None) == (not self._weights) or not self._scores, "weights must be given for all batches or none"
        self._scores.append(output.to(self.device, torch.float32))
        self._targets.append(target)

    def value(self):
        """-> FloatTensor [K] (CPU, like the reference) with the average precision of each class; 0 when empty."""
        if not self._scores:
            return 0
        scores = torch.cat(self._scores, 0).t().contiguous()          # [K,N]
        targets = torch.cat(self._targets, 0).t().contiguous()
        K, N = scores.shape
        _, ind = torch.sort(scores, dim=1, descending=True, stable=True)
        truth = torch.gather(targets, 1, ind).contiguous()
        wsorted = None
        if self._weights:
            wsorted = torch.cat(self._weights, 0)[ind].contiguous()   # [K,N]
        ap = torch.empty(K, device=self.device, dtype=torch.float32)
        call("cf_ap_sorted", ptr(truth), ptr(wsorted) if wsorted is not None else None, ptr(ap), N, K, stream_ptr())
        return ap.cpu()
