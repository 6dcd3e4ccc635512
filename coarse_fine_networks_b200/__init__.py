"""coarse_fine_networks_b200 -- B200-native (sm_100a) Coarse-Fine X3D hot path.

Host-side mirror of the reference's nn.Module surface (x3d_fine.py, x3d_coarse.py,
interp1d.py) over the C ABI in include/cfnet_b200.h.  Importing the package loads
libcfnet_b200.so and fails loudly if it has not been built; there is no CPU fallback."""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)

__all__ = ["_lib"]
