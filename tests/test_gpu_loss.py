"""GPU parity of the fused loss kernels against the reference goldens (values and gradients)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


def test_dice_and_focal_match_reference_golden():
    from segmentation3d.loss.binary_dice_loss import BinaryDiceLoss
    from segmentation3d.loss.focal_loss import FocalLoss
    from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
    z = np.load(os.path.join(G, 'loss.npz'))
    for c in (2, 5):
        k = 'c%d_' % c
        probs, target, w = torch.from_numpy(z[k + 'probs']).cuda(), torch.from_numpy(z[k + 'target']).cuda(), z[k + 'weights'].tolist()
        p = probs.clone().requires_grad_(True)
        l = MultiDiceLoss(w, c, True)(p, target)
        l.backward()
        assert l.dim() == 0 and abs(l.item() - float(z[k + 'dice'])) <= 2e-6
        assert np.abs(p.grad.cpu().numpy() - z[k + 'dice_grad']).max() <= 1e-7 + 1e-4 * np.abs(z[k + 'dice_grad']).max()
        p = probs.clone().requires_grad_(True)
        l = FocalLoss(c, alpha=w, gamma=2, size_average=True, use_gpu=True)(p, target)
        l.backward()
        assert abs(l.item() - float(z[k + 'focal'])) <= 2e-6
        assert np.abs(p.grad.cpu().numpy() - z[k + 'focal_grad']).max() <= 1e-7 + 1e-4 * np.abs(z[k + 'focal_grad']).max()
        p = probs.clone().requires_grad_(True)
        l = FocalLoss(c, alpha=None, gamma=0, size_average=False, use_gpu=True)(p, target)
        l.backward()
        assert abs(l.item() - float(z[k + 'focal_g0_sum'])) <= 1e-5 * abs(float(z[k + 'focal_g0_sum']))
        assert np.abs(p.grad.cpu().numpy() - z[k + 'focal_g0_sum_grad']).max() <= 1e-4 * np.abs(z[k + 'focal_g0_sum_grad']).max()
    # BinaryDiceLoss standalone (generic two-channel input -> tensor-op form)
    l = BinaryDiceLoss()(torch.from_numpy(z['bin_probs']).cuda(), torch.from_numpy(z['bin_target']).cuda())
    assert abs(l.item() - float(z['bin_dice'])) <= 1e-6


def test_dice_full_size_properties():
    """config-3 size (B=8, C=2, 96^3): loss of a perfect prediction is ~0, of the complement ~1, and the
    fused reduction equals the closed form evaluated with torch reductions."""
    from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
    g = torch.Generator(device='cuda').manual_seed(0)
    B, C, D = 8, 2, 96
    target = torch.randint(0, C, (B, 1, D, D, D), generator=g, device='cuda').float()
    onehot = torch.cat([(target == i).float() for i in range(C)], 1)
    loss = MultiDiceLoss([0.5, 0.5], C, True)
    assert loss(onehot, target).item() <= 1e-6
    assert abs(loss(1.0 - onehot, target).item() - 1.0) <= 1e-6
    probs = torch.softmax(torch.randn((B, C, D, D, D), generator=g, device='cuda'), 1)
    got = loss(probs, target).item()
    ref = 0.0
    for i in range(C):
        q = (probs[:, i] * (probs[:, i] > 1.0 / C)).double().flatten(1)
        t = (target[:, 0] == i).double().flatten(1)
        ref += 0.5 * float((1 - (2 * (q * t).sum(1) + 1e-6) / ((q * q).sum(1) + t.sum(1) + 1e-6)).mean())
    assert abs(got - ref) <= 1e-6


def test_cross_entropy_matches_reference_golden():
    """loss.name = 'CE' (core/seg_train.py:98-99): fused log-softmax + NLL kernels against values and gradients of the
    reference's wrapper - default, class weights, an ignored label, 'sum' and 'none' - plus the legacy flags."""
    from segmentation3d.loss.cross_entropy_loss import CrossEntropyLoss
    z, zc = np.load(os.path.join(G, 'loss.npz')), np.load(os.path.join(G, 'loss_ce.npz'))
    cases = (('ce', {}), ('ce_w', {'weight': True}), ('ce_ign', {'weight': True, 'ignore_index': 1}),
             ('ce_sum', {'reduction': 'sum'}), ('ce_none', {'weight': True, 'ignore_index': 0, 'reduction': 'none'}))
    for c in (2, 5):
        k = 'c%d_' % c
        probs, target = torch.from_numpy(z[k + 'probs']).cuda(), torch.from_numpy(z[k + 'target']).cuda()
        for tag, kw in cases:
            kw = dict(kw)
            if kw.pop('weight', False):
                kw['weight'] = torch.from_numpy(zc[k + 'weight'])
            p = probs.clone().requires_grad_(True)
            l = CrossEntropyLoss(**kw)(p, target)
            assert tuple(l.shape) == tuple(zc[k + tag].shape), (k, tag)
            (l if l.dim() == 0 else (l * torch.arange(l.numel(), dtype=torch.float32, device='cuda').view_as(l) / l.numel()).sum()).backward()
            ref, gref = zc[k + tag], zc[k + tag + '_grad']
            assert np.abs(l.detach().cpu().numpy() - ref).max() <= 1e-5 * max(1.0, float(np.abs(ref).max())), (k, tag)
            assert np.abs(p.grad.cpu().numpy() - gref).max() <= 1e-7 + 1e-4 * np.abs(gref).max(), (k, tag)
    assert CrossEntropyLoss(None, False).reduction == 'sum' and CrossEntropyLoss(None, None, -100, False).reduction == 'none'
    x2 = torch.randn((7, 3), generator=torch.Generator().manual_seed(1))
    t2 = torch.tensor([0, 2, 1, 1, 0, 2, 2])
    got = CrossEntropyLoss()(x2.cuda(), t2.cuda()).item()
    assert abs(got - float(torch.nn.functional.cross_entropy(x2, t2))) <= 1e-5
