"""Deterministic synthetic tensors / weights shared by the golden generator and the tests.

CPU ``torch.Generator`` streams are reproducible for a fixed torch build, and the build
container and the GPU box run the same image, so a (shape, seed) pair names the same tensor
on both sides; that keeps the committed fixtures small (outputs only)."""
import zlib

import torch


def synth_tensor(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(int(seed))
    return torch.randn(*shape, generator=g, dtype=torch.float32) * scale


def _fill(key, t, seed):
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    if not t.is_floating_point():
        return t
    if key.endswith("running_var"):
        return torch.rand(t.shape, generator=g) + 0.5
    if key.endswith("running_mean"):
        return torch.randn(t.shape, generator=g) * 0.1
    if t.dim() == 1 and key.endswith("weight"):
        return torch.rand(t.shape, generator=g) + 0.5
    if t.dim() == 1:
        return torch.randn(t.shape, generator=g) * 0.1
    fan_in = t[0].numel()
    return torch.randn(t.shape, generator=g) * (2.0 / fan_in) ** 0.5


def synth_state_dict(template, seed):
    """template: mapping key -> tensor (shapes/dtypes only are used)."""
    return {k: _fill(k, v, seed).to(v.dtype) for k, v in template.items()}


def fill_state_dict(module, seed):
    sd = module.state_dict()
    module.load_state_dict(synth_state_dict(sd, seed), strict=True)
    return module
