"""End-to-end drop-in flow on the GPU: seg_train.train -> model folder -> seg_infer.segmentation, checked against
the oracle run on the checkpoint the training wrote."""
import os

import numpy as np
import pytest
import torch

from oracle import sliding_window as osw
from oracle.metrics import parity_report

pytestmark = pytest.mark.gpu


def _write_cfg(path, text):
    with open(path, 'w') as f:
        f.write(text)


def test_train_then_infer_cli_flow(tmp_path, monkeypatch):
    monkeypatch.setenv('SEG3D_MODE', 'fp32')
    from segmentation3d.core.seg_infer import segmentation
    from segmentation3d.core.seg_train import train
    from segmentation3d.utils.image3d import Image3d, read_image, write_image
    rng = np.random.default_rng(0)
    data = tmp_path / 'data'
    os.makedirs(data)
    lines = ['2']
    for i in range(2):
        vol = (rng.standard_normal((48, 48, 64)) * 200).astype(np.float32)
        zz, yy, xx = np.mgrid[0:48, 0:48, 0:64]
        mask = (((zz - 24) ** 2 + (yy - 24) ** 2 + (xx - 30 - 4 * i) ** 2) < 15 ** 2).astype(np.int8)
        vol += mask.astype(np.float32) * 400
        write_image(Image3d(vol), str(data / ('im%d.mha' % i)), True)
        write_image(Image3d(mask), str(data / ('seg%d.mha' % i)), True)
        lines += [str(data / ('im%d.mha' % i)), str(data / ('seg%d.mha' % i))]
    _write_cfg(str(data / 'train.txt'), '\n'.join(lines) + '\n')
    save_dir = tmp_path / 'model'
    cfg = """
from easydict import EasyDict as edict
from segmentation3d.utils.normalizer import FixedNormalizer
__C = edict()
cfg = __C
__C.general = {}
__C.general.imseg_list = %r
__C.general.save_dir = %r
__C.general.model_scale = 'fine'
__C.general.resume_epoch = -1
__C.general.num_gpus = 1
__C.general.seed = 0
__C.dataset = {}
__C.dataset.num_classes = 2
__C.dataset.spacing = [1.0, 1.0, 1.0]
__C.dataset.crop_size = [32, 32, 32]
__C.dataset.sampling_method = 'HYBRID'
__C.dataset.interpolation = 'LINEAR'
__C.dataset.crop_normalizers = [FixedNormalizer(0.0, 400.0, True)]
__C.dataset.random_translation = [4, 4, 4]
__C.dataset.random_scale = [1.0, 1.0]
__C.loss = {}
__C.loss.name = 'Dice'
__C.loss.obj_weight = [0.5, 0.5]
__C.loss.focal_gamma = 2
__C.net = {}
__C.net.name = 'vnet'
__C.train = {}
__C.train.epochs = 3
__C.train.batchsize = 2
__C.train.num_threads = 0
__C.train.lr = 1e-3
__C.train.betas = (0.9, 0.999)
__C.train.save_epochs = 1
__C.debug = {}
__C.debug.save_inputs = False
""" % (str(data / 'train.txt'), str(save_dir))
    _write_cfg(str(tmp_path / 'train_config.py'), cfg)
    train(str(tmp_path / 'train_config.py'))
    chk_root = save_dir / 'fine' / 'checkpoints'
    assert sorted(os.listdir(chk_root)) == ['chk_1', 'chk_2']
    state = torch.load(str(chk_root / 'chk_2' / 'params.pth'), weights_only=False, map_location='cpu')
    assert state['net'] == 'vnet' and state['max_stride'] == 16 and state['out_channels'] == 2
    assert all(k.startswith('module.') for k in state['state_dict'])          # loadable by the reference GPU path
    assert 'train_loss' in open(str(save_dir / 'fine' / 'train_log.txt')).read()
    # the template infer_config.py was copied next to the model; switch it to single-scale sliding window
    icfg = open(str(save_dir / 'infer_config.py')).read()
    icfg += "\n__C.general.single_scale = 'fine'\n__C.fine.partition_size = [32, 32, 32]\n__C.fine.partition_stride = [16, 16, 16]\n__C.fine.pick_largest_cc = False\n"
    _write_cfg(str(save_dir / 'infer_config.py'), icfg)
    out = tmp_path / 'out'
    masks = segmentation(str(data / 'im0.mha'), str(save_dir), str(out), 'seg.mha', 0, True, True, True, True)
    case = out / 'im0.mha'
    assert sorted(os.listdir(case)) == ['mean_prob_0.mha', 'mean_prob_1.mha', 'org.mha', 'seg.mha']
    seg = read_image(str(case / 'seg.mha')).to_numpy()
    p1 = read_image(str(case / 'mean_prob_1.mha')).to_numpy()
    assert seg.dtype == np.int8 and np.array_equal(seg, masks[0].to_numpy())
    # oracle on the very checkpoint the training wrote
    vol = read_image(str(data / 'im0.mha'), np.float32).to_numpy()
    probs, mask, _, _ = osw.segmentation_volume(state['state_dict'], vol, [1.0, 1.0, 1.0], state['crop_normalizers'][0], 'SIZE',
                                                [32, 32, 32], [16, 16, 16], 16, double_forward=False, faithful_copies=False)
    rep = parity_report(probs, np.stack([1 - p1, p1], 0))
    print('e2e parity', rep)
    assert np.abs(probs[1] - p1).max() <= 1e-3
    assert (mask == seg).mean() >= 0.999
    # batch inference over a txt case list, sharded by case as under `torch.distributed.run` (BASELINE configs[4]):
    # rank 1 of 2 segments only the second case; nothing is exchanged between ranks
    _write_cfg(str(data / 'test.txt'), '2\ncase0 %s\ncase1 %s\n' % (data / 'im0.mha', data / 'im1.mha'))
    monkeypatch.setenv('WORLD_SIZE', '2')
    monkeypatch.setenv('RANK', '1')
    monkeypatch.setenv('LOCAL_RANK', '0')
    out2 = tmp_path / 'out_sharded'
    masks2 = segmentation(str(data / 'test.txt'), str(save_dir), str(out2), 'seg.mha', 0, True, True, False, False)
    assert sorted(os.listdir(out2)) == ['case1'] and len(masks2) == 1
    monkeypatch.setenv('RANK', '0')
    masks0 = segmentation(str(data / 'test.txt'), str(save_dir), str(out2), 'seg.mha', 0, True, True, False, False)
    assert sorted(os.listdir(out2)) == ['case0', 'case1'] and len(masks0) == 1
    assert np.array_equal(masks0[0].to_numpy(), seg)               # case0 = im0.mha: same mask as the single-file call
