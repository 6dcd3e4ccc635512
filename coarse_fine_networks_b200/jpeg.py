"""GPU JPEG decode for the loaders (cf_jpeg_* of the C ABI, nvJPEG batched decoder): the JPEG streams of a clip ->
ONE uint8 CUDA tensor [T,H,W,3], the input of spatial_transforms.Compose.clip().  Replaces the host decode of the
reference's pil_loader / video_loader / load_rgb_frames (charades_fine.py:22-27, 46-56, 78-101).

Decoded pixels are standard-conforming but not bit-identical to PIL's libjpeg-turbo (different IDCT / chroma up-sampling
arithmetic); tests/test_jpeg_gpu.py states the measured tolerance.  One decoder per host thread."""
import ctypes

import torch

from ._lib import call, lib, stream_ptr


class JpegDecoder:
    def __init__(self):
        self._h = ctypes.c_void_p()
        call("cf_jpeg_create", ctypes.cast(ctypes.byref(self._h), ctypes.c_void_p))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.cf_jpeg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def image_size(self, blob):
        """-> (H, W) of one JPEG stream (header parse only)."""
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        call("cf_jpeg_image_info", self._h, ctypes.cast(buf, ctypes.c_void_p), len(blob), ctypes.cast(ctypes.byref(h), ctypes.c_void_p),
             ctypes.cast(ctypes.byref(w), ctypes.c_void_p), ctypes.cast(ctypes.byref(c), ctypes.c_void_p))
        return h.value, w.value

    def decode(self, blobs, device="cuda", out=None):
        """blobs: list of `bytes` (JPEG streams of equally sized frames) -> uint8 CUDA tensor [n,H,W,3] (RGB).
        The call returns after the decode was enqueued on the current stream; the host buffers are kept alive by nvJPEG's
        own staging (nvjpegDecodeBatched consumes the bit streams before it returns)."""
        n = len(blobs)
        if n == 0:
            raise ValueError("decode(): no frames")
        H, W = self.image_size(blobs[0])
        dev = torch.device(device)
        if out is None:
            out = torch.empty(n, H, W, 3, device=dev, dtype=torch.uint8)
        elif not (out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (n, H, W, 3) and out.is_contiguous()):
            raise RuntimeError(f"decode(): out must be a contiguous uint8 CUDA tensor [{n},{H},{W},3]")
        bufs = [(ctypes.c_ubyte * len(b)).from_buffer_copy(b) for b in blobs]
        ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p).value for b in bufs])
        lens = (ctypes.c_size_t * n)(*[len(b) for b in blobs])
        with torch.cuda.device(out.device):
            call("cf_jpeg_decode_batch", self._h, ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(lens, ctypes.c_void_p), n,
                 out.data_ptr(), H, W, stream_ptr())
        return out
