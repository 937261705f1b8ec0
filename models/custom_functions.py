from ar_nerf_b200.custom_functions import (RayAABBIntersector, RayMarcher, RaySphereIntersector, TruncExp,  # noqa: F401
                                           VolumeRenderer)
