"""Build-container fuzz of the oracle against the UNMODIFIED reference beyond the committed golden fixtures: losses (values
and gradients) over random class counts / weights / shapes, the Dice-ratio metric, the network forward on random small
shapes with randomised GroupNorm affines, and seg_infer.segmentation_volume on random volumes / partitions / normalisers.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/fuzz_oracle_vs_reference.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

sitk = ref_shims.install()
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
from oracle import init as oinit, loss as oloss, metrics as ometrics, net as onet, sliding_window as osw   # noqa: E402
sys.path.insert(0, '/root/reference')
from easydict import EasyDict as edict                                          # noqa: E402
from segmentation3d.core import seg_infer as ref_infer                          # noqa: E402
from segmentation3d.loss.binary_dice_loss import BinaryDiceLoss                 # noqa: E402
from segmentation3d.loss.cross_entropy_loss import CrossEntropyLoss             # noqa: E402
from segmentation3d.loss.focal_loss import FocalLoss                            # noqa: E402
from segmentation3d.loss.multi_dice_loss import MultiDiceLoss                   # noqa: E402
from segmentation3d.network import vbnet as ref_vbnet, vnet as ref_vnet         # noqa: E402
from segmentation3d.utils.metrics import cal_dsc as ref_dsc                     # noqa: E402
from segmentation3d.utils.normalizer import AdaptiveNormalizer, FixedNormalizer  # noqa: E402

rng = np.random.default_rng(11)
g = torch.Generator().manual_seed(11)
worst = {}


def note(name, err):
    worst[name] = max(worst.get(name, 0.0), float(err))


# ---- losses -------------------------------------------------------------------------------------------------------------
for it in range(60):
    c = int(rng.integers(2, 7))
    shape = (int(rng.integers(1, 4)), c, int(rng.integers(2, 7)), int(rng.integers(2, 7)), int(rng.integers(2, 9)))
    probs = torch.softmax(torch.randn(shape, generator=g) * float(rng.choice([0.1, 1.0, 4.0])), 1)
    if rng.random() < 0.5:
        probs[:, :, 0] = 1.0 / c                                  # exact ties
    target = torch.randint(0, c, (shape[0], 1) + shape[2:], generator=g).float()
    w = [float(v) for v in rng.uniform(0.2, 3.0, c)]
    gamma = float(rng.choice([0, 1, 2, 3]))
    sa = bool(rng.random() < 0.5)
    for name, ref_fn, ora_fn in (
            ('dice', lambda p: MultiDiceLoss(w, c, False)(p, target), lambda p: oloss.multi_dice_loss(p, target, w)),
            ('focal', lambda p: FocalLoss(c, alpha=w, gamma=gamma, size_average=sa, use_gpu=False)(p, target),
             lambda p: oloss.focal_loss(p, target, c, alpha=w, gamma=gamma, size_average=sa)),
            ('ce', lambda p: CrossEntropyLoss()(p, target), lambda p: oloss.cross_entropy_loss(p, target))):
        p1, p2 = probs.clone().requires_grad_(True), probs.clone().requires_grad_(True)
        l1, l2 = ref_fn(p1), ora_fn(p2)
        l1.backward()
        l2.backward()
        note(name + ' value', abs(float(l1) - float(l2)) / max(1.0, abs(float(l1))))
        note(name + ' grad', float((p1.grad - p2.grad).abs().max()) / max(1e-12, float(p1.grad.abs().max())))
    if c == 2:
        note('binary dice', abs(float(BinaryDiceLoss()(probs.clone(), target)) - float(oloss.binary_dice_loss(probs.clone(), target))))

# ---- Dice-ratio metric ------------------------------------------------------------------------------------------------------
for it in range(200):
    a = rng.integers(0, 4, size=(6, 7, 8))
    b = a.copy()
    b[rng.random(a.shape) < rng.choice([0.0, 0.05, 0.5, 1.0])] = 0
    label, thr = int(rng.integers(0, 4)), int(rng.choice([1, 10, 100, 400]))
    r, o = ref_dsc(a, b, label, thr), ometrics.cal_dsc(a, b, label, thr)
    assert r[1] == o[1], (r, o)
    note('cal_dsc', abs(float(r[0]) - float(o[0])))

# ---- network forward ------------------------------------------------------------------------------------------------------------
for arch, mod in (('vnet', ref_vnet), ('vbnet', ref_vbnet)):
    for it in range(3):
        cout, seed = int(rng.integers(2, 6)), int(rng.integers(0, 1000))
        torch.manual_seed(seed)
        net = mod.SegmentationNet(1, cout)
        mod.parameters_kaiming_init(net)
        sd = oinit.randomize_affine({k: v.detach().clone() for k, v in net.state_dict().items()}, seed + 1)
        net.load_state_dict(sd)
        net.eval()
        shape = (int(rng.integers(1, 3)), 1, 16 * int(rng.integers(1, 3)), 16 * int(rng.integers(1, 3)), 16 * int(rng.integers(1, 4)))
        x = torch.randn(shape, generator=g)
        with torch.no_grad():
            ref = net(x)
        note(arch + ' forward', float((ref - onet.forward(sd, x)).abs().max()))
        ini = oinit.init_state_dict(arch, 1, cout, seed)
        torch.manual_seed(seed)
        net2 = mod.SegmentationNet(1, cout)
        mod.parameters_kaiming_init(net2)
        assert all(torch.equal(ini[k], v) for k, v in net2.state_dict().items()), 'seeded init differs'

# ---- sliding window -----------------------------------------------------------------------------------------------------------------
for it in range(4):
    size = [16 * int(rng.integers(2, 5)) for _ in range(3)]
    psize = [float(rng.choice([32, 48])) for _ in range(3)]
    pstride = [float(ps * rng.choice([0.5, 1.0])) for ps in psize]
    adaptive = bool(rng.random() < 0.5)
    torch.manual_seed(it)
    net = ref_vnet.SegmentationNet(1, 2)
    ref_vnet.parameters_kaiming_init(net)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net.eval()
    model = edict()
    model.net = net
    model.spacing, model.max_stride, model.interpolation = [1.0, 1.0, 1.0], 16, 'LINEAR'
    model.in_channels, model.out_channels = 1, 2
    model.crop_normalizers = [AdaptiveNormalizer(2.0)] if adaptive else [FixedNormalizer(10.0, 200.0, True)]
    cfg = edict()
    cfg.partition_type, cfg.partition_size, cfg.partition_stride = 'SIZE', psize, pstride
    cfg.cpu_model_spacing_increase_ratio, cfg.cpu_partition_decrease_ratio = 1.0, 1.0
    cfg.pick_largest_cc, cfg.remove_small_cc = False, 0
    vol = (rng.standard_normal((size[2], size[1], size[0])) * 250).astype(np.float32)
    mean_probs, mask = ref_infer.segmentation_volume(model, cfg, sitk.GetImageFromArray(vol), None, None, False)
    p_ref = np.stack([sitk.GetArrayFromImage(p) for p in mean_probs], 0)
    nd = {'type': 1, 'clip_sigma': 2.0} if adaptive else {'type': 0, 'mean': 10.0, 'stddev': 200.0, 'clip': True}
    p, m, _, _ = osw.segmentation_volume(sd, vol, [1.0, 1.0, 1.0], nd, 'SIZE', psize, pstride, 16)
    note('segmentation_volume probs', float(np.abs(p - p_ref).max()))
    note('segmentation_volume mask mismatch', float((m != sitk.GetArrayFromImage(mask).astype(np.int8)).mean()))

for k in sorted(worst):
    print('%-36s worst difference %.3g' % (k, worst[k]))
