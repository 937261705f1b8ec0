"""`import vren` drop-in: the reference's pybind11 extension name, served by ar_nerf_b200.vren (libarnerf.so)."""
from ar_nerf_b200.vren import *  # noqa: F401,F403
from ar_nerf_b200.vren import (composite_test_fw, composite_train_bw, composite_train_fw, distortion_loss_bw,  # noqa: F401
                               distortion_loss_fw, morton3D, morton3D_invert, packbits, ray_aabb_intersect,
                               ray_sphere_intersect, raymarching_test, raymarching_train)
