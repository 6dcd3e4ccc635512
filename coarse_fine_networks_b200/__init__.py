"""coarse_fine_networks_b200 -- B200-native (sm_100a) Coarse-Fine X3D hot path.

Host-side mirror of the reference's nn.Module surface (x3d_fine.py, x3d_coarse.py,
interp1d.py) over the C ABI in include/cfnet_b200.h.  Every op module imports ``_lib``,
which loads libcfnet_b200.so and raises if it has not been built
(``python -m coarse_fine_networks_b200.build``); there is no CPU / PyTorch fallback."""
