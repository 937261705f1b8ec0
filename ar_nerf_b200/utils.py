"""Checkpoint helpers of the reference (utils.py:4-39), same names and behaviour: a checkpoint is either a plain state dict
or a pytorch-lightning file whose 'state_dict' holds the model's entries under the prefix `model.` (train.py:53-60 names the
NGP `self.model`).  The released weights (README.md:79) load into ar_nerf_b200.networks.NGP unchanged: same state-dict keys
(xyz_encoder.params, rgb_net.params, center, xyz_min, xyz_max, half_size, density_bitfield, density_grid, grid_coords,
tonemapper_net_{0,1,2}.params), both hash-grid geometries tiny-cuda-nn builds have produced (NGP._load_from_state_dict)."""
import torch


def extract_model_state_dict(ckpt_path, model_name='model', prefixes_to_ignore=[]):
    """utils.py:4-19."""
    checkpoint = torch.load(ckpt_path, map_location='cpu', weights_only=False)
    checkpoint_ = {}
    if 'state_dict' in checkpoint:  # a pytorch-lightning checkpoint
        checkpoint = checkpoint['state_dict']
    for k, v in checkpoint.items():
        if not k.startswith(model_name):
            continue
        k = k[len(model_name) + 1:]
        if any(k.startswith(prefix) for prefix in prefixes_to_ignore):
            continue
        checkpoint_[k] = v
    return checkpoint_


def load_ckpt(model, ckpt_path, model_name='model', prefixes_to_ignore=[]):
    """utils.py:22-27.  Entries the checkpoint lacks keep the model's values; buffers train.py registers on the fly
    (density_grid, grid_coords: train.py:79-82) are created first when the checkpoint carries them."""
    if not ckpt_path:
        return
    checkpoint_ = extract_model_state_dict(ckpt_path, model_name, prefixes_to_ignore)
    if ('density_grid' in checkpoint_ or 'grid_coords' in checkpoint_) and hasattr(model, 'init_density_grid'):
        model.init_density_grid()
    model_dict = model.state_dict()
    model_dict.update(checkpoint_)
    model.load_state_dict(model_dict)


def slim_ckpt(ckpt_path, save_poses=False):
    """utils.py:30-39: the lightning checkpoint's state dict without what inference does not need."""
    ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=False)
    keys_to_pop = ['directions', 'model.density_grid', 'model.grid_coords']
    if not save_poses:
        keys_to_pop += ['poses']
    for k in ckpt['state_dict']:
        if k.startswith('val_lpips'):
            keys_to_pop += [k]
    for k in keys_to_pop:
        ckpt['state_dict'].pop(k, None)
    return ckpt['state_dict']
