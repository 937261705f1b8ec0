from ar_nerf_b200.rendering import *  # noqa: F401,F403
from ar_nerf_b200.rendering import MAX_SAMPLES, NEAR_DISTANCE, render, render_surface_normal, render_surface_rgb  # noqa: F401
