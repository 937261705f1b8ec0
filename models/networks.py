from ar_nerf_b200.networks import NGP  # noqa: F401
