"""Clip input pipeline on CPU: the numpy oracle (oracle/clip_oracle.py) against the outputs of the
reference's own transform classes + Pillow (tests/golden/clip_pipeline.npz, made by tests/golden/make_golden.py), the
library's HOST coefficient function against the oracle, and the mirror classes' random draws against the reference's."""
import os
import random

import numpy as np
import pytest

from oracle import clip_oracle as CO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "clip_pipeline.npz")
SCALES = [224 / 256., 224 / 320.]


@pytest.fixture(scope="module")
def gold():
    d = np.load(GOLD)
    return {k: d[k] for k in d.files}


def case_params(gold, name):
    frames = gold[f"{name}/frames"]
    size, seed, train = (int(v) for v in gold[f"{name}/meta"])
    H, W = frames.shape[1:3]
    if train:
        scale, tlx, tly, p = gold[f"{name}/draw"]
        return frames, CO.multiscale_crop_box(W, H, scale, tlx, tly), size, bool(p < 0.5), seed, True
    return frames, CO.center_crop_box(W, H), size, False, seed, False


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_oracle_bit_exact_vs_reference_transforms(gold):
    for name in gold["names"]:
        frames, box, size, flip, _, _ = case_params(gold, str(name))
        ours = CO.clip_preprocess(frames, box, size, flip, gold["mean"], gold["std"])
        ref = gold[f"{name}/clip"]
        assert ours.shape == ref.shape, name
        assert np.array_equal(bits(ours), bits(ref)), f"{name}: max diff {np.abs(ours - ref).max()}"


def test_oracle_resample_vs_pillow_ramps(gold):
    for key in [k for k in gold if k.startswith("ramp/")]:
        i, o = (int(v) for v in key[5:].split("_"))
        ramp = (np.arange(i * 3) * 37 % 256).astype(np.uint8).reshape(1, i, 3)
        b, k = CO.resample_coeffs(i, o)
        assert np.array_equal(CO._pass_u8(ramp, b, k, 1)[0], gold[key]), key


def test_zero_padding_like_collate(gold):
    frames, box, size, flip, _, _ = case_params(gold, "train_up_112")
    out = CO.clip_preprocess(frames, box, size, flip, gold["mean"], gold["std"], t_pad=5)
    assert out.shape == (3, 5, size, size)
    assert np.array_equal(bits(out[:, :3]), bits(gold["train_up_112/clip"])) and not out[:, 3:].any()


def test_host_coefficient_tables_match_oracle():
    """cf_resample_ksize / cf_resample_coeffs are host functions of the C ABI (no GPU needed)."""
    from coarse_fine_networks_b200 import _lib
    for (i, o) in [(210, 224), (168, 224), (240, 224), (256, 224), (224, 224), (105, 112), (84, 112), (120, 112), (270, 64),
                   (157, 160), (315, 312), (360, 312), (1080, 224), (3, 8), (1000, 4)]:
        rb, rk = CO.resample_coeffs(i, o)
        ks = _lib.lib.cf_resample_ksize(i, o)
        assert ks == rk.shape[1]
        b, k = np.zeros((o, 2), np.int32), np.zeros((o, ks), np.int32)
        assert _lib.lib.cf_resample_coeffs(i, o, b.ctypes.data, k.ctypes.data) == 0
        assert np.array_equal(b, rb) and np.array_equal(k, rk), (i, o)
    assert _lib.lib.cf_resample_coeffs(0, 8, None, None) != 0 and b"bad arguments" in _lib.lib.cf_last_error()


def test_mirror_classes_draw_like_the_reference(gold):
    """Same python `random` seed -> same scale / offsets / flip as the reference's classes drew (stored in the golden)."""
    from coarse_fine_networks_b200 import spatial_transforms as ST
    for name in gold["names"]:
        name = str(name)
        frames, box, size, flip, seed, train = case_params(gold, name)
        H, W = frames.shape[1:3]
        if train:
            tr = ST.Compose([ST.MultiScaleRandomCropMultigrid(SCALES, size), ST.RandomHorizontalFlip(), ST.ToTensor(255),
                             ST.Normalize(list(gold["mean"]), list(gold["std"]))])
        else:
            tr = ST.Compose([ST.CenterCropScaled(size), ST.ToTensor(255), ST.Normalize(list(gold["mean"]), list(gold["std"]))])
        random.seed(seed)
        tr.randomize_parameters(size)
        assert tr.params(W, H) == (box[0], box[1], box[2], size, flip), name


def test_compose_rejects_chains_the_scripts_never_build():
    from coarse_fine_networks_b200 import spatial_transforms as ST
    with pytest.raises(NotImplementedError):
        ST.Compose([ST.ToTensor(255), ST.Normalize([0, 0, 0], [1, 1, 1])])
    with pytest.raises(NotImplementedError):
        ST.ToTensor(1)
    tr = ST.Compose([ST.CenterCropScaled(8), ST.ToTensor(255), ST.Normalize([0, 0, 0], [1, 1, 1])])
    import torch
    with pytest.raises(RuntimeError):                      # no CPU fallback
        tr.clip(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))


def test_clip_entry_point_rejects_bad_arguments_before_touching_the_gpu():
    """Argument validation of cf_clip_preprocess happens on the host, ahead of any CUDA call, so it can be exercised here:
    non-multiple-of-4 size, crop box outside the frame, coefficient tables built for another crop/size pair, short channel stride."""
    import ctypes
    from coarse_fine_networks_b200 import _lib
    f = _lib.lib.cf_clip_preprocess
    p = ctypes.c_void_p(0x1000)                                # never dereferenced: every call below fails validation first

    def call(T=2, H=40, W=60, x1=0, y1=0, crop=40, size=32, ks=None, rows_max=30, flip=0, t_out=2, stride=None):
        ks = _lib.lib.cf_resample_ksize(crop, size) if ks is None else ks
        stride = t_out * size * size if stride is None else stride
        rc = f(p, p, p, p, p, p, p, T, H, W, x1, y1, crop, size, ks, ks, rows_max, flip, t_out, stride, None)
        return rc, _lib.lib.cf_last_error().decode()

    for kw, msg in ((dict(size=30), "multiple of 4"), (dict(x1=30), "crop box outside"), (dict(y1=1), "crop box outside"),
                    (dict(ks=7), "do not match"), (dict(stride=100), "channel stride"), (dict(t_out=1), "T <= t_out"),
                    (dict(rows_max=0), "rows_max")):
        rc, err = call(**kw)
        assert rc != 0 and msg in err, (kw, rc, err)
