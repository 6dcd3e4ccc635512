"""CPU oracle of the clip input pipeline (TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product).

Restates, in numpy integer / fp32 arithmetic, what the reference's loader does to every decoded RGB frame
(SURVEY 8(f) next-4):

    charades_fine.py:170-172         spatial_transform.randomize_parameters(224); [transform(img) for img in imgs];
                                     torch.stack(...).permute(1, 0, 2, 3)               -> [3,T,S,S]
    transforms/spatial_transforms.py 488-503  MultiScaleRandomCropMultigrid: crop box + img.resize(BILINEAR)
                                     216-230  CenterCropScaled (validation):  centre box + img.resize(BILINEAR)
                                     342-354  RandomHorizontalFlip
                                     46-87    ToTensor(255): uint8 HWC -> float CHW / 255
                                     108-118  Normalize(mean, std): t.sub_(m).div_(s)
    charades_fine.py:215-226         mt_collate_fn: zero padding of the shorter clips of a batch

`img.resize` is Pillow's ImagingResample (third-party: Pillow 12.2.0, src/libImaging/Resample.c, not vendored in
the reference): a separable triangle filter whose support grows with the down-scale factor, coefficients computed
in double, normalised, rounded to 22-bit fixed point (PRECISION_BITS = 32 - 8 - 2), horizontal pass first, each pass
rounded and clipped to uint8.  Pinned: tests/golden/clip_pipeline.npz holds outputs of the reference's own transform
classes (and therefore of Pillow) produced by tests/golden/make_golden.py; tests/test_clip_cpu.py checks every
function here bit-exactly against them.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter over the
    whole axis [0, in_size) -> out_size.  Returns (bounds int32 [out,2] = (first tap, tap count), kk int32 [out,ksize])."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0.0 else a
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] = w[:xmax] / ww
        for x in range(xmax):
            v = w[x] * float(1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass_u8(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One Pillow 8-bit resampling pass along `axis` (0 = vertical, 1 = horizontal) of an [H,W,C] uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for j in range(n):
            acc += src[lo + j] * int(kk[o, j])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img: np.ndarray, size: int) -> np.ndarray:
    """`PIL.Image.resize((size, size), BILINEAR)` of an [H,W,3] uint8 image (horizontal pass, then vertical)."""
    h, w = img.shape[:2]
    bh, kh = resample_coeffs(w, size)
    bv, kv = resample_coeffs(h, size)
    return _pass_u8(_pass_u8(img, bh, kh, 1), bv, kv, 0)


def multiscale_crop_box(w: int, h: int, scale: float, tl_x: float, tl_y: float):
    """spatial_transforms.py:490-500 -> (x1, y1, crop_size)."""
    crop = int(min(w, h) * scale)
    return int(tl_x * (w - crop)), int(tl_y * (h - crop)), crop


def center_crop_box(w: int, h: int):
    """spatial_transforms.py:222-225 -> (x1, y1, crop_size).  Python round(): half to even."""
    crop = min(w, h)
    return int(round((w - crop) / 2.0)), int(round((h - crop) / 2.0)), crop


def normalize_lut(mean, std) -> np.ndarray:
    """[3,256] fp32: ToTensor(255) then Normalize -- ((v / 255) - m) / s, every step rounded to fp32."""
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    return np.stack([(v - np.float32(m)) / np.float32(s) for m, s in zip(mean, std)]).astype(np.float32)


def clip_preprocess(frames: np.ndarray, box, size: int, flip: bool, mean, std, t_pad: int = 0) -> np.ndarray:
    """frames [T,H,W,3] uint8 -> [3, max(T,t_pad), size, size] fp32 (crop, resize, flip, /255, normalise, zero pad)."""
    x1, y1, crop = box
    lut = normalize_lut(mean, std)
    T = frames.shape[0]
    out = np.zeros((3, max(T, t_pad), size, size), np.float32)
    for t in range(T):
        r = resize_bilinear_u8(frames[t, y1:y1 + crop, x1:x1 + crop], size)
        if flip:
            r = r[:, ::-1]
        for c in range(3):
            out[c, t] = lut[c][r[..., c]]
    return out
