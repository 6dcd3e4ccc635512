"""GPU clip input pipeline with the surface of the reference's transforms/spatial_transforms.py for the
transforms its training / validation loaders compose (train_fine.py:74-80):

    Compose([MultiScaleRandomCropMultigrid(scales, size), RandomHorizontalFlip(), ToTensor(255), Normalize(m, s)])
    Compose([CenterCropScaled(size), ToTensor(255), Normalize(m, s)])

The classes keep the reference's constructor arguments, `randomize_parameters(c_size, index)` protocol and draw
order from Python's `random` (spatial_transforms.py:32-34, 505-509, 356-357), so the same seed selects the same
crops and flips.  The reference then calls the pipeline once per PIL image on the CPU and stacks / permutes /
zero-pads (charades_fine.py:170-172, 215-226); here `Compose.clip(frames)` takes all decoded frames of a video
as one uint8 CUDA tensor [T,H,W,3] and produces the normalised [3,T,S,S] clip with ONE kernel
(cf_clip_preprocess: crop + Pillow-exact bilinear resize + flip + /255 + normalise + padding), bit-identical to
the reference's output.  `collate_clips` fills a [B,3,Tmax,S,S] batch in place, one launch per video.

JPEG decoding is not part of this module (frames arrive decoded, e.g. from nvJPEG or a host decoder).  Transforms
the shipped scripts never compose (Scale, CenterCrop, CornerCrop, MultiScaleCornerCrop, MultiScaleRandomCrop,
RandomVerticalFlip, CenterCropScaledMultiple) are not provided.  No CPU fallback: tensors must be on a CUDA device.
"""
import random

import numpy as np
import torch

from ._lib import call, lib, ptr, stream_ptr

_BAND = 16          # CLIP_BAND of csrc/clip_input.cu


class _Tables:
    """Device copies of Pillow's coefficient tables, cached per (crop, size, device)."""
    _cache = {}

    @classmethod
    def get(cls, crop, size, device):
        key = (int(crop), int(size), str(device))
        hit = cls._cache.get(key)
        if hit is not None:
            return hit
        ks = lib.cf_resample_ksize(crop, size)
        bounds = np.zeros((size, 2), np.int32)
        kk = np.zeros((size, ks), np.int32)
        rc = lib.cf_resample_coeffs(crop, size, bounds.ctypes.data, kk.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"cf_resample_coeffs failed ({rc}): {lib.cf_last_error().decode()}")
        rows_max = 1
        for oy0 in range(0, size, _BAND):
            last = min(oy0 + _BAND, size) - 1
            rows_max = max(rows_max, int(bounds[last, 0] + bounds[last, 1] - bounds[oy0, 0]))
        entry = (torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device), ks, rows_max)
        cls._cache[key] = entry
        return entry


class ToTensor:
    """spatial_transforms.py:37-90.  Only the uint8 RGB path (norm_value applied as a float division)."""

    def __init__(self, norm_value=255):
        if norm_value != 255:
            raise NotImplementedError("ToTensor: only norm_value=255 (train_fine.py:76,79) is built")
        self.norm_value = norm_value

    def randomize_parameters(self, c_size=0, index=0):
        pass


class Normalize:
    """spatial_transforms.py:93-121: channel = (channel - mean) / std."""

    def __init__(self, mean, std):
        assert len(mean) == 3 and len(std) == 3, "RGB mean / std expected"
        self.mean, self.std = list(mean), list(std)

    def randomize_parameters(self, c_size=0, index=0):
        pass


class RandomHorizontalFlip:
    """spatial_transforms.py:339-357: flips when the drawn p < 0.5."""
    p = 1.0

    def randomize_parameters(self, c_size=0, index=0):
        self.p = random.random()


class MultiScaleRandomCropMultigrid:
    """spatial_transforms.py:480-509: square crop of side int(min(w,h)*scale) at a random corner offset, resized
    to (size, size) with PIL BILINEAR."""

    def __init__(self, scales, size, interpolation=None):
        self.scales = scales
        self.init_size = size
        self.size = self.init_size
        self.scale, self.tl_x, self.tl_y = scales[0], 0.0, 0.0

    def randomize_parameters(self, c_size, index=0):
        self.size = c_size
        self.scale = self.scales[random.randint(0, len(self.scales) - 1)]
        self.tl_x = random.random()
        self.tl_y = random.random()

    def box(self, w, h):
        crop = int(min(w, h) * self.scale)
        return int(self.tl_x * (w - crop)), int(self.tl_y * (h - crop)), crop


class CenterCropScaled:
    """spatial_transforms.py:201-233: centre square of side min(w,h), resized to `size` with PIL BILINEAR."""

    def __init__(self, size, interpolation=None):
        self.size = int(size) if not isinstance(size, (tuple, list)) else int(size[0])
        if isinstance(size, (tuple, list)) and size[0] != size[1]:
            raise NotImplementedError("CenterCropScaled: square output only")

    def randomize_parameters(self, c_size=0, index=0):
        pass

    def box(self, w, h):
        crop = min(w, h)
        return int(round((w - crop) / 2.)), int(round((h - crop) / 2.)), crop


class Compose:
    """spatial_transforms.py:18-34.  The composed chain is lowered onto one kernel; it must be
    [crop transform, (RandomHorizontalFlip), ToTensor(255), Normalize] as in train_fine.py:74-80."""

    def __init__(self, transforms):
        self.transforms = list(transforms)
        kinds = [type(t) for t in self.transforms]
        if not self.transforms or kinds[0] not in (MultiScaleRandomCropMultigrid, CenterCropScaled):
            raise NotImplementedError("Compose: the first transform must be MultiScaleRandomCropMultigrid or CenterCropScaled")
        rest = kinds[1:]
        if rest not in ([RandomHorizontalFlip, ToTensor, Normalize], [ToTensor, Normalize]):
            raise NotImplementedError("Compose: expected [crop, (RandomHorizontalFlip), ToTensor(255), Normalize] (train_fine.py:74-80)")
        self._crop = self.transforms[0]
        self._flip = self.transforms[1] if rest[0] is RandomHorizontalFlip else None
        self._norm = self.transforms[-1]
        self._lut = {}

    def randomize_parameters(self, c_size=0, index=0):
        for t in self.transforms:
            t.randomize_parameters(c_size, index)

    def get_state(self):
        """The current random draw as plain numbers: a loader WORKER draws (same `random` order as the reference) and ships
        this with the undecoded frames; the main process restores it with set_state() before clip()."""
        c, f = self._crop, self._flip
        st = {"size": int(c.size)}
        if isinstance(c, MultiScaleRandomCropMultigrid):
            st.update(scale=float(c.scale), tl_x=float(c.tl_x), tl_y=float(c.tl_y))
        if f is not None:
            st["flip_p"] = float(f.p)
        return st

    def set_state(self, st):
        c, f = self._crop, self._flip
        if isinstance(c, MultiScaleRandomCropMultigrid):
            c.size, c.scale, c.tl_x, c.tl_y = st["size"], st["scale"], st["tl_x"], st["tl_y"]
        if f is not None:
            f.p = st["flip_p"]

    # ------------------------------------------------------------------------------------------------------
    def _lut_on(self, device):
        key = str(device)
        if key not in self._lut:
            lut = torch.empty(3, 256, device=device, dtype=torch.float32)
            m, s = self._norm.mean, self._norm.std
            call("cf_normalize_lut", ptr(lut), float(m[0]), float(m[1]), float(m[2]), float(s[0]), float(s[1]), float(s[2]), stream_ptr())
            self._lut[key] = lut
        return self._lut[key]

    def params(self, w, h):
        """-> (x1, y1, crop, size, flip) for frames of width w and height h with the current random draw."""
        x1, y1, crop = self._crop.box(w, h)
        flip = self._flip is not None and self._flip.p < 0.5
        return x1, y1, crop, int(self._crop.size), bool(flip)

    def clip(self, frames, out=None, t_pad=0):
        """frames: uint8 CUDA tensor [T,H,W,3] (all frames of one video share one random draw, charades_fine.py:170-171).
        -> fp32 [3, max(T,t_pad), S, S]; frames beyond T are zeros.  `out` may be a [3,Tout,S,S] view into a batch."""
        if not (torch.is_tensor(frames) and frames.is_cuda and frames.dtype == torch.uint8):
            raise RuntimeError("clip(): frames must be a uint8 CUDA tensor [T,H,W,3] (no CPU fallback)")
        if frames.dim() != 4 or frames.shape[3] != 3:
            raise RuntimeError(f"clip(): expected [T,H,W,3], got {tuple(frames.shape)}")
        frames = frames.contiguous()
        T, H, W, _ = frames.shape
        x1, y1, crop, S, flip = self.params(W, H)
        t_out = max(T, int(t_pad))
        if out is None:
            out = torch.empty(3, t_out, S, S, device=frames.device, dtype=torch.float32)
        else:
            t_out = out.shape[1]
            ok = (out.is_cuda and out.dtype == torch.float32 and out.dim() == 4 and out.shape[0] == 3 and t_out >= T
                  and tuple(out.shape[2:]) == (S, S) and out.stride(3) == 1 and out.stride(2) == S and out.stride(1) == S * S)
            if not ok:
                raise RuntimeError("clip(): out must be fp32 CUDA [3,Tout>=T,S,S] with dense frames")
        with torch.cuda.device(frames.device):
            bounds, kk, ks, rows_max = _Tables.get(crop, S, frames.device)
            lut = self._lut_on(frames.device)
            call("cf_clip_preprocess", ptr(frames) if T else None, ptr(out), ptr(bounds), ptr(kk), ptr(bounds), ptr(kk), ptr(lut),
                 T, H, W, x1, y1, crop, S, ks, ks, rows_max, int(flip), t_out, out.stride(0), stream_ptr())
        return out

    def __call__(self, img):
        """One frame: uint8 CUDA tensor [H,W,3] -> fp32 [3,S,S] (the reference's per-image call)."""
        return self.clip(img.unsqueeze(0))[:, 0]


def collate_clips(videos, transform, c_size=224, randomize=True):
    """The clip half of Charades.__getitem__ + mt_collate_fn (charades_fine.py:169-172, 215-226): one random draw per
    video, every video resampled into its slot of a zero-padded [B,3,Tmax,S,S] batch.  videos: list of uint8 CUDA
    tensors [T_i,H_i,W_i,3] (frame sizes may differ between videos).  -> (batch, lengths list)."""
    assert len(videos) > 0
    t_max = max(int(v.shape[0]) for v in videos)
    batch = None
    for b, v in enumerate(videos):
        if randomize:
            transform.randomize_parameters(c_size)
        if batch is None:
            S = transform.params(int(v.shape[2]), int(v.shape[1]))[3]
            batch = torch.empty(len(videos), 3, t_max, S, S, device=v.device, dtype=torch.float32)
        transform.clip(v, out=batch[b])
    return batch, [int(v.shape[0]) for v in videos]
