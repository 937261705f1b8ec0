"""Checkpoint helpers with the reference's names and behaviour (utils.py:4-39).  A checkpoint file is either a plain state
dict or a pytorch-lightning file whose 'state_dict' holds the model's entries under `model.` (train.py:53-60 names the NGP
`self.model`).  The released weights (README.md:79) load into ar_nerf_b200.networks.NGP unchanged: same state-dict keys
(xyz_encoder.params, rgb_net.params, center, xyz_min, xyz_max, half_size, density_bitfield, density_grid, grid_coords,
tonemapper_net_{0,1,2}.params), both hash-grid geometries tiny-cuda-nn builds have produced (NGP._load_from_state_dict)."""
import torch


def _read(path):
    return torch.load(path, map_location='cpu', weights_only=False)


def extract_model_state_dict(ckpt_path, model_name='model', prefixes_to_ignore=[]):
    """utils.py:4-19: the entries whose key starts with `model_name`, that name and its separator cut off, minus the entries whose
    remaining key starts with one of `prefixes_to_ignore`."""
    blob = _read(ckpt_path)
    entries = blob['state_dict'] if 'state_dict' in blob else blob   # lightning file | plain state dict
    cut = len(model_name) + 1
    skip = tuple(prefixes_to_ignore)
    return {name[cut:]: value for name, value in entries.items()
            if name.startswith(model_name) and not name[cut:].startswith(skip)}


def load_ckpt(model, ckpt_path, model_name='model', prefixes_to_ignore=[]):
    """utils.py:22-27: what the checkpoint lacks keeps the model's value.  The buffers train.py registers on the fly
    (density_grid, grid_coords: train.py:79-82) are created first when the checkpoint carries them."""
    if not ckpt_path:
        return
    found = extract_model_state_dict(ckpt_path, model_name, prefixes_to_ignore)
    if hasattr(model, 'init_density_grid') and not {'density_grid', 'grid_coords'}.isdisjoint(found):
        model.init_density_grid()
    model.load_state_dict({**model.state_dict(), **found})


def slim_ckpt(ckpt_path, save_poses=False):
    """utils.py:30-39: the lightning checkpoint's state dict without what inference does not need (pixel directions, the
    occupancy grid's float form and coordinates, the LPIPS network, and -- unless asked for -- the optimised poses)."""
    state = _read(ckpt_path)['state_dict']
    unused = {'directions', 'model.density_grid', 'model.grid_coords'}
    if not save_poses:
        unused.add('poses')
    unused.update(name for name in state if name.startswith('val_lpips'))
    for name in unused:
        state.pop(name, None)
    return state
