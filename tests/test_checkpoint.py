"""Checkpoint path: ar_nerf_b200.utils.load_ckpt / slim_ckpt / extract_model_state_dict with the reference's
semantics (utils.py:4-39) on files laid out the way train.py's LightningModule writes them (keys under `model.` next to
`directions`, `poses`, `val_lpips.*`), round-tripped through ar_nerf_b200.networks.NGP."""
import os

import torch

from ar_nerf_b200.networks import NGP
from ar_nerf_b200.utils import extract_model_state_dict, load_ckpt, slim_ckpt


def _lightning_ckpt(model, path, extra=True):
    sd = {"model." + k: v.clone() for k, v in model.state_dict().items()}
    if extra:
        sd.update({"directions": torch.zeros(4, 3), "poses": torch.zeros(2, 3, 4), "val_lpips.net.weight": torch.zeros(3),
                   "dR": torch.zeros(2, 3), "dT": torch.zeros(2, 3)})
    torch.save({"state_dict": sd, "epoch": 3, "global_step": 3000}, path)


def test_round_trip_lightning_checkpoint(tmp_path):
    torch.manual_seed(0)
    src = NGP(0.5)
    src.init_density_grid()
    with torch.no_grad():
        src.xyz_encoder.params.normal_(); src.rgb_net.params.normal_()
        src.density_grid.uniform_(-1, 5); src.density_bitfield.random_(0, 255)
    path = os.path.join(tmp_path, "epoch=29.ckpt")
    _lightning_ckpt(src, path)
    assert set(extract_model_state_dict(path)) == set(src.state_dict())
    assert "density_grid" not in extract_model_state_dict(path, prefixes_to_ignore=["density_grid", "grid_coords"])
    dst = NGP(0.5)
    load_ckpt(dst, path)  # the checkpoint carries density_grid / grid_coords: the buffers are created as train.py:79-82 does
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    # slim_ckpt drops what inference does not need and still loads (utils.py:30-39, train.py:318-323)
    slim = slim_ckpt(path)
    assert not any(k in slim for k in ("directions", "poses", "model.density_grid", "model.grid_coords", "val_lpips.net.weight"))
    assert "poses" in slim_ckpt(path, save_poses=True)
    slim_path = os.path.join(tmp_path, "slim.ckpt")
    torch.save(slim, slim_path)                # a plain state dict, as train.py:323 saves it
    dst2 = NGP(0.5)
    load_ckpt(dst2, slim_path)
    assert torch.equal(dst2.xyz_encoder.params, src.xyz_encoder.params) and torch.equal(dst2.density_bitfield, src.density_bitfield)
    assert not hasattr(dst2, "density_grid")   # slimmed away, and load_ckpt did not invent it
    load_ckpt(dst2, "")                        # utils.py:23: an empty path is a no-op


def test_hdr_checkpoint_keys_and_partial_load(tmp_path):
    src = NGP(0.5, rgb_act='None')
    keys = set(src.state_dict())
    assert {f"tonemapper_net_{i}.params" for i in range(3)} <= keys
    path = os.path.join(tmp_path, "hdr.ckpt")
    _lightning_ckpt(src, path, extra=False)
    dst = NGP(0.5, rgb_act='None')
    before = dst.rgb_net.params.clone()
    load_ckpt(dst, path, prefixes_to_ignore=["rgb_net"])  # ignored prefixes keep the model's own values (utils.py:13-17)
    assert torch.equal(dst.rgb_net.params, before)
    assert torch.equal(dst.tonemapper_net_1.params, src.tonemapper_net_1.params)
    # the tonemapper reads its 1-wide input padded with ones, as tiny-cuda-nn does: the 15 padded columns of W1 act as a bias
    x = torch.linspace(-2, 2, 7)[:, None]
    with torch.no_grad():
        y0 = dst.tonemapper_net_0(x)
        dst.tonemapper_net_0.params[1:16] += 0.5   # row 0 of W1, padded columns
        y1 = dst.tonemapper_net_0(x)
    assert y0.shape == (7, 1) and not torch.equal(y0, y1)
