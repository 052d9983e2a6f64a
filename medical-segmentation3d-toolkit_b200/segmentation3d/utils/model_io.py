"""Checkpoint folder layout (drop-in for reference utils/model_io.py:7-93):

    <save_dir>/<model_scale>/checkpoints/chk_<epoch>/{params.pth, optimizer.pth, train_config.py}

params.pth keys: epoch, batch, net, max_stride, state_dict, spacing, interpolation, in_channels,
out_channels, crop_normalizers.  Files written here load in the reference and vice versa.
"""
import glob
import os
import shutil

import torch


def get_checkpoint_folder(chk_root, epoch):
    """Folder of checkpoint `epoch`; epoch < 0 picks the largest chk_<int> (model_io.py:7-28)."""
    assert os.path.isdir(chk_root), 'The folder does not exist: {}'.format(chk_root)
    if epoch < 0:
        found = [int(os.path.basename(p).split('_')[-1]) for p in glob.glob(os.path.join(chk_root, 'chk_*'))]
        epoch = max(found) if found else -1
    return os.path.join(chk_root, 'chk_{}'.format(epoch))


def load_checkpoint(epoch_idx, net, opt, save_dir):
    """Restore net + optimizer; returns (epoch, batch) (model_io.py:31-54)."""
    folder = os.path.join(save_dir, 'checkpoints', 'chk_{}'.format(epoch_idx))
    chk_file = os.path.join(folder, 'params.pth')
    assert os.path.isfile(chk_file), 'checkpoint file not found: {}'.format(chk_file)
    state = torch.load(chk_file, map_location='cpu', weights_only=False)
    net.load_state_dict(state['state_dict'])
    opt_file = os.path.join(folder, 'optimizer.pth')
    assert os.path.isfile(opt_file), 'optimizer file not found: {}'.format(opt_file)
    opt.load_state_dict(torch.load(opt_file, map_location='cpu', weights_only=False))
    return state['epoch'], state['batch']


def save_checkpoint(net, opt, epoch_idx, batch_idx, cfg, max_stride, num_modality):
    """Write params.pth / optimizer.pth / train_config.py (model_io.py:57-93)."""
    model_folder = os.path.join(cfg.general.save_dir, cfg.general.model_scale)
    chk_folder = os.path.join(model_folder, 'checkpoints', 'chk_{}'.format(epoch_idx))
    os.makedirs(chk_folder, exist_ok=True)
    state = {
        'epoch': epoch_idx, 'batch': batch_idx, 'net': cfg.net.name, 'max_stride': max_stride,
        'state_dict': net.state_dict(), 'spacing': cfg.dataset.spacing, 'interpolation': cfg.dataset.interpolation,
        'in_channels': num_modality, 'out_channels': cfg.dataset.num_classes,
        'crop_normalizers': [n.to_dict() for n in cfg.dataset.crop_normalizers],
    }
    torch.save(state, os.path.join(chk_folder, 'params.pth'))
    torch.save(opt.state_dict(), os.path.join(chk_folder, 'optimizer.pth'))
    src = os.path.join(model_folder, 'train_config.py')
    if os.path.isfile(src):
        shutil.copy(src, os.path.join(chk_folder, 'train_config.py'))
