"""`from utils import load_ckpt, slim_ckpt` drop-in (the reference's top-level utils.py), served by ar_nerf_b200.utils."""
from ar_nerf_b200.utils import extract_model_state_dict, load_ckpt, slim_ckpt  # noqa: F401
