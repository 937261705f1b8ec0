"""ar_nerf_b200 -- B200 (sm_100a) implementation of the AR-NeRF / ngp_pl rendering hot path.

Host-side mirror of the reference's interface for this path:
    ar_nerf_b200.rendering.render            <- models/rendering.py:14
    ar_nerf_b200.networks.NGP                <- models/networks.py:12
    ar_nerf_b200.custom_functions.*          <- models/custom_functions.py
    ar_nerf_b200.losses.{NeRFLoss,DistortionLoss}  <- losses.py
    ar_nerf_b200.vren.*                      <- the pybind module `vren` (models/csrc/binding.cpp:234-250)
The top-level `models/` package and `vren.py` re-export them under the reference's import paths.
All arithmetic runs in libarnerf.so (include/arnerf.h); there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["rendering", "networks", "custom_functions", "losses", "vren", "field", "trainer", "workload"]
