"""Clip input pipeline on the GPU (cf_clip_preprocess through spatial_transforms.Compose): bit-exact against the goldens
made by the reference's own transforms + Pillow, against the numpy oracle on seeded inputs (ragged sizes, both chains,
large down-scale), zero padding / in-place batch collation, and the size-independent properties at full size."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import clip_oracle as CO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "clip_pipeline.npz")
MEAN, STD = [0.413, 0.368, 0.338], [0.131, 0.125, 0.132]
SCALES = [224 / 256., 224 / 320.]


@pytest.fixture(scope="module")
def ST():
    import __graft_entry__ as ge
    ge.build()
    from coarse_fine_networks_b200 import spatial_transforms
    return spatial_transforms


def chain(ST, train, size):
    if train:
        return ST.Compose([ST.MultiScaleRandomCropMultigrid(SCALES, size), ST.RandomHorizontalFlip(), ST.ToTensor(255),
                           ST.Normalize(MEAN, STD)])
    return ST.Compose([ST.CenterCropScaled(size), ST.ToTensor(255), ST.Normalize(MEAN, STD)])


def same_bits(a, b):
    return torch.equal(a.detach().cpu().contiguous().view(torch.int32), torch.from_numpy(np.ascontiguousarray(b)).view(torch.int32))


def test_golden_reference_transforms_bit_exact(ST):
    d = np.load(GOLD)
    for name in d["names"]:
        name = str(name)
        size, seed, train = (int(v) for v in d[f"{name}/meta"])
        tr = chain(ST, train, size)
        random.seed(seed)
        tr.randomize_parameters(size)                       # the reference's call protocol (charades_fine.py:170)
        out = tr.clip(torch.from_numpy(d[f"{name}/frames"]).cuda())
        assert same_bits(out, d[f"{name}/clip"]), f"{name}: {(out.cpu() - torch.from_numpy(d[f'{name}/clip'])).abs().max()}"
        one = tr(torch.from_numpy(d[f"{name}/frames"][0]).cuda())              # per-image call
        assert same_bits(one, d[f"{name}/clip"][:, 0])


@pytest.mark.parametrize("T,H,W,size,train,seed", [(3, 97, 131, 64, True, 3), (2, 97, 131, 64, False, 0), (2, 131, 97, 100, True, 4),
                                                   (1, 480, 640, 32, False, 0), (2, 60, 80, 224, True, 9), (1, 33, 33, 32, False, 0)])
def test_vs_oracle_ragged(ST, T, H, W, size, train, seed):
    rng = np.random.default_rng(seed + 50)
    frames = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    tr = chain(ST, train, size)
    random.seed(seed)
    tr.randomize_parameters(size)
    x1, y1, crop, S, flip = tr.params(W, H)
    ref = CO.clip_preprocess(frames, (x1, y1, crop), S, flip, MEAN, STD, t_pad=T + 2)
    out = tr.clip(torch.from_numpy(frames).cuda(), t_pad=T + 2)
    assert out.shape == ref.shape and same_bits(out, ref)
    assert not out[:, T:].any()


def test_collate_into_batch_in_place(ST):
    """Videos of different length and frame size into one zero-padded batch (charades_fine.py:215-226)."""
    rng = np.random.default_rng(7)
    vids = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((4, 72, 96, 3), (2, 90, 120, 3), (5, 64, 64, 3))]
    tr = chain(ST, True, 64)
    random.seed(11)
    batch, lens = ST.collate_clips([torch.from_numpy(v).cuda() for v in vids], tr, c_size=64)
    assert batch.shape == (3, 3, 5, 64, 64) and lens == [4, 2, 5]
    random.seed(11)
    for b, v in enumerate(vids):
        tr.randomize_parameters(64)
        x1, y1, crop, S, flip = tr.params(v.shape[2], v.shape[1])
        assert same_bits(batch[b], CO.clip_preprocess(v, (x1, y1, crop), S, flip, MEAN, STD, t_pad=5)), b


def test_full_size_properties(ST):
    """cfg-size clip [64 frames, 240x320 -> 224]: (i) flip == mirror of the unflipped result; (ii) a constant frame maps to
    the normalised constant; (iii) identity crop (size == crop) reproduces the normalisation table of the cropped pixels."""
    rng = np.random.default_rng(1)
    frames = torch.from_numpy(rng.integers(0, 256, (64, 240, 320, 3), dtype=np.uint8)).cuda()
    tr = chain(ST, True, 224)
    random.seed(2)
    tr.randomize_parameters(224)
    tr.transforms[1].p = 0.9
    a = tr.clip(frames)
    tr.transforms[1].p = 0.1
    b = tr.clip(frames)
    assert torch.equal(a.flip(-1), b)
    lut = torch.from_numpy(CO.normalize_lut(MEAN, STD)).cuda()
    const = torch.full((2, 240, 320, 3), 77, dtype=torch.uint8, device="cuda")
    c = tr.clip(const)
    assert all(torch.equal(c[ch], torch.full_like(c[ch], float(lut[ch, 77]))) for ch in range(3))
    tv = chain(ST, False, 240)                              # centre crop 240 of 240x320 -> 240: every tap weight is 1
    v = tv.clip(frames[:3])
    crop = frames[:3, :, 40:280].long()
    assert all(torch.equal(v[ch], lut[ch][crop[..., ch]]) for ch in range(3))


def test_argument_errors(ST):
    tr = chain(ST, False, 30)                               # size % 4 != 0
    with pytest.raises(RuntimeError, match="multiple of 4"):
        tr.clip(torch.zeros(1, 40, 40, 3, dtype=torch.uint8, device="cuda"))
    tr = chain(ST, False, 32)
    with pytest.raises(RuntimeError):
        tr.clip(torch.zeros(1, 40, 40, 4, dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError):
        tr.clip(torch.zeros(1, 40, 40, 3, dtype=torch.float32, device="cuda"))
