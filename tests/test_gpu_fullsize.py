"""Full-size checks (BASELINE configs[1]): 96^3 patches in fp16 against the oracle, and size-independent
properties of the whole 512x512x400 sliding-window pass."""
import numpy as np
import pytest
import torch

from oracle import init as oinit
from oracle import net as onet
from oracle import sliding_window as osw
from oracle.metrics import parity_report

pytestmark = pytest.mark.gpu


def _model(mode, batch=6):
    from segmentation3d.core.seg_infer import make_model
    from segmentation3d.network import vnet
    net = vnet.SegmentationNet(1, 2)
    net.load_state_dict(oinit.init_state_dict('vnet', 1, 2, 0))
    net.b200_mode = mode
    net = net.cuda().eval()
    return make_model(net, [1.0, 1.0, 1.0], {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True}, batch=batch)


def _ct(size_xyz, seed):
    g = torch.Generator().manual_seed(seed)
    X, Y, Z = size_xyz
    lo = torch.randn((1, 1, max(2, Z // 32), max(2, Y // 32), max(2, X // 32)), generator=g)
    field = torch.nn.functional.interpolate(lo, size=(Z, Y, X), mode='trilinear', align_corners=False)[0, 0]
    return (field * 600.0 + torch.randn((Z, Y, X), generator=g) * 60.0 - 200.0).clamp_(-1000.0, 2000.0).float()


def test_vnet_96_patch_fp16_meets_bars():
    sd = oinit.init_state_dict('vnet', 1, 2, 0)
    g = torch.Generator().manual_seed(3)
    x = torch.randn((1, 1, 96, 96, 96), generator=g)
    ref = onet.forward(sd, x)
    from segmentation3d._b200.plan import NetPlan
    y = NetPlan(sd, mode='fp16', device='cuda').forward(x.cuda()).cpu()
    rep = parity_report(ref[0].numpy(), y[0].numpy())
    print('VNet 96^3 fp16', rep)
    assert rep['max_abs'] <= 1e-2 and rep['agree'] >= 0.999 and min(rep['dice']) >= 0.999, rep


def test_config2_full_volume_properties():
    from segmentation3d.core.seg_infer import segmentation_volume_device
    model = _model('fp16')
    cfg = {'partition_type': 'SIZE', 'partition_size': [96, 96, 96], 'partition_stride': [96, 96, 96]}
    vol = _ct([512, 512, 400], 1234).cuda()
    acc, mask = segmentation_volume_device(model, cfg, vol)
    torch.cuda.synchronize()
    assert acc.shape == (2, 400, 512, 512) and mask.shape == (400, 512, 512) and mask.dtype == torch.int8
    # probabilities: normalised by the overlap count, so every voxel sums to 1 and lies in [0, 1]
    s = acc.sum(0)
    assert float((s - 1).abs().max()) <= 1e-3
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-6
    # mask is the first argmax of the returned probabilities (ties -> class 0)
    assert torch.equal(mask, acc.max(0)[1].to(torch.int8))
    # idempotence
    acc2, mask2 = segmentation_volume_device(model, cfg, vol)
    assert torch.equal(mask, mask2) and float((acc - acc2).abs().max()) <= 1e-6
    # patch sharding (rank::world) followed by a sum equals the single pass
    eng = model['engine']
    starts, ends = osw.partition_grid([512, 512, 400], [1, 1, 1], [0, 0, 0], [512, 512, 400], [96] * 3, [96] * 3, 16)
    assert len(starts) == 180
    from segmentation3d._b200.sliding import axis_counts
    parts = []
    for r in range(2):
        a = torch.zeros_like(acc)
        eng.accumulate(vol, starts[r::2], [96, 96, 96], {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True}, a)
        parts.append(a)
    tot = parts[0] + parts[1]
    m3 = eng.finalize(tot, axis_counts([512, 512, 400], starts, ends))
    assert float((tot - acc).abs().max()) <= 1e-5
    assert float((m3 == mask).float().mean()) >= 0.99999
    # eight patches drawn from the 180 (seeded; corner, face and interior patches occur) against the oracle forward of the same
    # crops: BASELINE.json reduced-precision bars per patch - max|dp| <= 1e-2, label agreement >= 99.9 %, per-class Dice >= 0.999
    sd = oinit.init_state_dict('vnet', 1, 2, 0)
    # only patches no other patch overlaps: 512 = 5 x 96 + 32 and 400 = 4 x 96 + 16, so the clamped last box of every axis
    # overlaps its neighbour and those two average their probabilities (checked against the reference blend elsewhere)
    alone = [i for i, s0 in enumerate(starts) if s0[0] <= 288 and s0[1] <= 288 and s0[2] <= 192]
    assert len(alone) == 4 * 4 * 3
    rng = np.random.default_rng(5)
    picked = sorted(rng.choice(alone, size=8, replace=False).tolist())
    worst = {'max_abs': 0.0, 'agree': 1.0, 'dice': 1.0}
    for i in picked:
        s0, e0 = starts[i], ends[i]
        crop = vol[s0[2]:e0[2], s0[1]:e0[1], s0[0]:e0[0]].cpu().numpy()
        ref = onet.forward(sd, torch.from_numpy(osw.normalize_fixed(crop, 0.0, 1000.0, True))[None, None])[0].numpy()
        got = acc[:, s0[2]:e0[2], s0[1]:e0[1], s0[0]:e0[0]].cpu().numpy()
        rep = parity_report(ref, got)
        print('config-2 patch %3d at %s vs oracle' % (i, s0), rep)
        worst = {'max_abs': max(worst['max_abs'], rep['max_abs']), 'agree': min(worst['agree'], rep['agree']),
                 'dice': min(worst['dice'], min(rep['dice']))}
        assert rep['max_abs'] <= 1e-2 and rep['agree'] >= 0.999 and min(rep['dice']) >= 0.999, (i, rep)
    print('config-2 worst over 8 patches', worst)
