"""Evaluation path on CPU: the oracle restatement of APMeter.value() and of the 25-point sampling
against the reference's own outputs (tests/golden/apmeter.npz, produced by /root/reference/apmeter.py)."""
import os

import numpy as np
import torch

from oracle import cf_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def test_average_precision_oracle_vs_reference():
    g = load("apmeter")
    ap = O.average_precision(g["scores"], g["targets"])
    assert torch.allclose(ap, g["ap"], rtol=0, atol=1e-6), (ap - g["ap"]).abs().max()
    apw = O.average_precision(g["scores"], g["targets"], g["weights"])
    assert torch.allclose(apw, g["ap_weighted"], rtol=0, atol=1e-6), (apw - g["ap_weighted"]).abs().max()
    assert float(g["ap"][5]) == 0.0                        # the class without positives


def test_localize_samples_matches_the_script_slicing():
    C, TL, valid = 4, 140, 131
    probs = torch.arange(C * TL, dtype=torch.float32).view(C, TL)
    labels = (probs % 3 == 0).float()
    p1, l1 = O.localize_samples(probs, labels, valid)
    sc = valid / 25.0
    assert torch.equal(p1, probs[:, :valid][:, 1::int(sc)][:, :25]) and p1.shape[1] == 25
    assert torch.equal(l1, labels[:, :valid][:, 1::int(sc)][:, :25])
