"""GPU parity of the standalone network modules (segmentation3d/network/module/*.py) against the oracle's restatement of
the reference blocks (the host wiring alone is also pinned on the CPU by tests/test_blocks_wiring.py)."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import net as onet

pytestmark = [pytest.mark.gpu]
EPS = 1e-5


def _sd(module, prefix=''):
    return {prefix + k: v.detach().cpu() for k, v in module.state_dict().items()}


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith('gn.weight') or name.endswith('gn1.weight') or name.endswith('gn2.weight'):
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
            elif name.endswith('.bias'):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            else:
                fan = p[0].numel() if p.dim() == 5 else 1
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / fan) ** 0.5)


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('fp16', 2e-2)])
def test_standalone_modules_match_oracle_blocks(monkeypatch, mode, tol):
    monkeypatch.setenv('SEG3D_MODE', mode)
    from segmentation3d.network.module.conv_gn_relu3 import ConvGnRelu3
    from segmentation3d.network.module.residual_block3 import BottResidualBlock3, ResidualBlock3
    from segmentation3d.network.module.vnet_downblock import DownBlock
    from segmentation3d.network.module.vnet_inblock import InputBlock
    from segmentation3d.network.module.vnet_outblock import OutputBlock
    from segmentation3d.network.module.vnet_upblock import UpBlock
    g = torch.Generator().manual_seed(0)

    def close(got, ref, what):
        got = got.cpu()
        assert got.shape == ref.shape and got.dtype == torch.float32, what
        err = float((got - ref).abs().max()) / max(1.0, float(ref.abs().max()))
        print(what, mode, 'max rel err %.3g' % err)
        assert err <= tol, (what, mode, err)

    with torch.no_grad():
        m = ConvGnRelu3(16, 32, 3, 1, 1, do_act=False)
        _randomize(m, 1)
        x = torch.randn((2, 16, 8, 8, 16), generator=g)
        close(m.cuda()(x.cuda()), F.group_norm(F.conv3d(x, m.conv.weight.cpu(), m.conv.bias.cpu(), padding=1), 1, m.gn.weight.cpu(), m.gn.bias.cpu(), EPS),
              'ConvGnRelu3')
        for blk, C in ((ResidualBlock3(32, 3, 1, 1, 2), 32), (BottResidualBlock3(64, 3, 1, 1, 4, 2), 64)):
            _randomize(blk, 3)
            x = torch.randn((2, C, 8, 8, 16), generator=g)
            close(blk.cuda()(x.cuda()), onet._rblock(x, _sd(blk, 'r.'), 'r'), type(blk).__name__)
        m = InputBlock(1, 16)
        _randomize(m, 4)
        x = torch.randn((2, 1, 16, 16, 16), generator=g)
        sd = _sd(m)
        close(m.cuda()(x.cuda()), F.relu(F.group_norm(F.conv3d(x, sd['conv.weight'], sd['conv.bias'], padding=1), 1, sd['gn.weight'], sd['gn.bias'], EPS)),
              'InputBlock')
        for comp in (False, True):
            m = DownBlock(32, 2, compression=comp)
            _randomize(m, 5)
            x = torch.randn((1, 32, 16, 16, 16), generator=g)
            close(m.cuda()(x.cuda()), onet._down(x, _sd(m, 'd.'), 'd'), 'DownBlock')
            m = UpBlock(128, 64, 2, compression=comp)
            _randomize(m, 6)
            x, skip = torch.randn((1, 128, 4, 4, 8), generator=g), torch.randn((1, 32, 8, 8, 16), generator=g)
            close(m.cuda()(x.cuda(), skip.cuda()), onet._up(x, skip, _sd(m, 'u.'), 'u'), 'UpBlock')
        for nc in (2, 5):
            m = OutputBlock(32, nc)
            _randomize(m, 7)
            x = torch.randn((2, 32, 8, 8, 16), generator=g)
            sd = _sd(m)
            y = F.relu(F.group_norm(F.conv3d(x, sd['conv1.weight'], sd['conv1.bias'], padding=1), 1, sd['gn1.weight'], sd['gn1.bias'], EPS))
            y = F.softmax(F.group_norm(F.conv3d(y, sd['conv2.weight'], sd['conv2.bias']), 1, sd['gn2.weight'], sd['gn2.bias'], EPS), 1)
            close(m.cuda()(x.cuda()), y, 'OutputBlock')


def test_segmentation_voi_and_disabled_partition_match_oracle():
    """core/seg_infer.py:208-246 (one VOI) and partition_type = 'DISABLE' (:277-279, the whole volume as one patch)."""
    import numpy as np
    from oracle import init as oinit
    from oracle import sliding_window as osw
    from segmentation3d.core.seg_infer import make_model, segmentation_voi, segmentation_volume
    from segmentation3d.network import vnet
    from segmentation3d.utils.image3d import Image3d
    sd = oinit.randomize_affine(oinit.init_state_dict('vnet', 1, 2, 0), 3)
    net = vnet.SegmentationNet(1, 2)
    net.load_state_dict(sd)
    net.b200_mode = 'fp32'
    nd = {'type': 0, 'mean': 10.0, 'stddev': 150.0, 'clip': True}
    model = make_model(net.cuda().eval(), [1.0, 1.0, 1.0], nd)
    vol = (torch.randn((48, 32, 64), generator=torch.Generator().manual_seed(2)) * 200).numpy().astype(np.float32)
    probs, mask, _, _ = osw.segmentation_volume(sd, vol, [1.0, 1.0, 1.0], nd, 'DISABLE', double_forward=False, faithful_copies=False)
    cfg = {'partition_type': 'DISABLE', 'pick_largest_cc': False, 'remove_small_cc': 0}
    p_im, m_im = segmentation_volume(model, cfg, Image3d(vol), None, None, True)
    got = np.stack([p.to_numpy() for p in p_im], 0)
    assert np.abs(got - probs).max() <= 1e-3 and (m_im.to_numpy() == mask).mean() >= 0.999
    maps = segmentation_voi(model, Image3d(vol, (1.0, 1.0, 1.0), (5.0, 6.0, 7.0)), [16, 0, 16], [48, 32, 48], True)
    ref, _, _, _ = osw.segmentation_volume(sd, vol[16:48, 0:32, 16:48], [1.0, 1.0, 1.0], nd, 'DISABLE', double_forward=False,
                                           faithful_copies=False)
    assert np.abs(np.stack([m.to_numpy() for m in maps], 0) - ref).max() <= 1e-3
    assert np.allclose(maps[0].GetOrigin(), (21.0, 6.0, 23.0)) and maps[0].GetSize() == (32, 32, 32)


def test_device_crops_match_host_crops(tmp_path):
    """seg3d_crop_resample against the host crop_image (ITK restatement) on the same volume: crops overlapping every face
    of the volume, rescaled, linear and nearest; then a DeviceCropLoader batch against the DataLoader batch."""
    import numpy as np
    from segmentation3d.utils.image3d import Image3d
    from segmentation3d.utils.image_tools import crop_image
    rng = np.random.default_rng(4)
    src = rng.standard_normal((40, 48, 56)).astype(np.float32)
    spacing, origin = (0.8, 1.1, 1.5), (-5.0, 3.0, 12.0)
    host = Image3d(src, spacing, origin)
    dev = Image3d(torch.from_numpy(src).cuda(), spacing, origin)
    for center in ((10.0, 20.0, 30.0), (-4.0, 3.5, 12.2), (39.0, 55.0, 70.0), (17.3, 28.9, 41.4)):
        for csp in ((0.8, 1.1, 1.5), (1.0, 1.0, 1.0), (0.55, 0.7, 2.1)):
            for interp in ('LINEAR', 'NN'):
                a = crop_image(host, center, [32, 24, 16], csp, interp)
                b = crop_image(dev, center, [32, 24, 16], csp, interp)
                assert b.is_cuda() and np.allclose(a.GetOrigin(), b.GetOrigin()) and np.allclose(a.GetSpacing(), b.GetSpacing())
                err = float(np.abs(a.to_numpy() - b.to_numpy()).max())
                assert err <= (1e-5 if interp == 'LINEAR' else 0.0), (center, csp, interp, err)


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-3), ('fp16', 1e-2)])
def test_baseline_config0_single_patch_cli_flow(tmp_path, monkeypatch, mode, tol):
    """BASELINE.json configs[0] literally: VNet random-init (seed 0, kaiming), checkpoint written by save_checkpoint into
    <model>/fine/checkpoints/chk_1/params.pth, infer_config with single_scale='fine' and partition_type='DISABLE', one
    synthetic 1x1x96^3 .mha patch through `segmentation()` - against the oracle on the same weights and volume."""
    import numpy as np
    from oracle import init as oinit
    from oracle import sliding_window as osw
    from oracle.metrics import parity_report
    from segmentation3d.core.seg_infer import segmentation
    from segmentation3d.network import vnet
    from segmentation3d.utils.attrdict import AttrDict
    from segmentation3d.utils.image3d import Image3d, read_image, write_image
    from segmentation3d.utils.model_io import save_checkpoint
    from segmentation3d.utils.normalizer import FixedNormalizer
    monkeypatch.setenv('SEG3D_MODE', mode)
    torch.manual_seed(0)
    net = vnet.SegmentationNet(1, 2)
    vnet.parameters_kaiming_init(net)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    assert all(torch.equal(sd[k], v) for k, v in oinit.init_state_dict('vnet', 1, 2, 0).items())
    model_dir = tmp_path / 'model'
    cfg = AttrDict({'general': {'save_dir': str(model_dir), 'model_scale': 'fine'}, 'net': {'name': 'vnet'},
                    'dataset': {'spacing': [1.0, 1.0, 1.0], 'interpolation': 'LINEAR', 'num_classes': 2,
                                'crop_normalizers': [FixedNormalizer(0.0, 1.0, False)]}})
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    save_checkpoint(net, opt, 1, 1, cfg, 16, 1)
    with open(str(model_dir / 'infer_config.py'), 'w') as f:
        f.write("from easydict import EasyDict as edict\n__C = edict()\ncfg = __C\n__C.general = {}\n__C.general.single_scale = 'fine'\n"
                "__C.fine = {}\n__C.fine.model_name = 'fine'\n__C.fine.pick_largest_cc = False\n__C.fine.remove_small_cc = 0\n"
                "__C.fine.partition_type = 'DISABLE'\n__C.fine.partition_size = [96, 96, 96]\n__C.fine.partition_stride = [96, 96, 96]\n"
                "__C.fine.cpu_model_spacing_increase_ratio = 1.0\n__C.fine.cpu_partition_decrease_ratio = 1.0\n")
    vol = torch.randn((96, 96, 96), generator=torch.Generator().manual_seed(9)).numpy().astype(np.float32)
    write_image(Image3d(vol), str(tmp_path / 'patch.mha'), False)
    out = tmp_path / 'out'
    masks = segmentation(str(tmp_path / 'patch.mha'), str(model_dir), str(out), 'seg.mha', 0, True, True, False, True)
    probs, mask, _, _ = osw.segmentation_volume(sd, vol, [1.0, 1.0, 1.0], {'type': 0, 'mean': 0.0, 'stddev': 1.0, 'clip': False},
                                                'DISABLE', double_forward=False, faithful_copies=False)
    got = np.stack([read_image(str(out / 'patch.mha' / ('mean_prob_%d.mha' % c))).to_numpy() for c in range(2)], 0)
    rep = parity_report(probs, got)
    print('configs[0]', mode, rep)
    assert rep['max_abs'] <= tol and rep['agree'] >= 0.999 and min(rep['dice']) >= 0.999
    assert np.array_equal(read_image(str(out / 'patch.mha' / 'seg.mha')).to_numpy(), masks[0].to_numpy())


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('fp16', 2e-2)])
def test_conv_gn_relu3_output_size_like_the_reference_test(monkeypatch, mode, tol):
    """network/module/conv_gn_relu3_test.py:8-43 on the GPU: k3 s1 p1 and k2 s2 p0 units from 1 to 16 channels on a
    [4, 1, 48, 32, 16] batch - the reference's shape assertions plus the values against torch's functional ops."""
    monkeypatch.setenv('SEG3D_MODE', mode)
    from segmentation3d.network.module.conv_gn_relu3 import ConvGnRelu3
    model1 = ConvGnRelu3(1, 16, ksize=3, stride=1, padding=1, do_act=True).cuda()
    model2 = ConvGnRelu3(1, 16, ksize=2, stride=2, padding=0, do_act=True).cuda()
    inputs = torch.rand([4, 1, 48, 32, 16], generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        outputs1, outputs2 = model1(inputs.cuda()).cpu(), model2(inputs.cuda()).cpu()
        sd1, sd2 = _sd(model1), _sd(model2)
        ref1 = F.relu(F.group_norm(F.conv3d(inputs, sd1['conv.weight'], sd1['conv.bias'], padding=1), 1, sd1['gn.weight'], sd1['gn.bias'], EPS))
        ref2 = F.relu(F.group_norm(F.conv3d(inputs, sd2['conv.weight'], sd2['conv.bias'], stride=2), 1, sd2['gn.weight'], sd2['gn.bias'], EPS))
    assert tuple(outputs1.size()) == (4, 16, 48, 32, 16) and tuple(outputs2.size()) == (4, 16, 24, 16, 8)
    assert float((outputs1 - ref1).abs().max()) <= tol * max(1.0, float(ref1.abs().max()))
    assert float((outputs2 - ref2).abs().max()) <= tol * max(1.0, float(ref2.abs().max()))
