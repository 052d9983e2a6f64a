"""GPU parity of the training step (forward with saved activations, backward kernels, Adam) against the
reference golden trace and the oracle's autograd."""
import os

import numpy as np
import pytest
import torch

from oracle import init as oinit
from oracle import loss as oloss
from oracle import net as onet

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


def _net(arch, cout, sd, mode):
    import importlib
    mod = importlib.import_module('segmentation3d.network.' + arch)
    net = mod.SegmentationNet(1, cout)
    net.load_state_dict(sd)
    net.b200_mode = mode
    return net.cuda().train()


@pytest.mark.parametrize('arch,cout,lossname', [('vnet', 2, 'dice'), ('vbnet', 5, 'focal')])
def test_gradients_match_oracle_autograd_fp32(arch, cout, lossname):
    from segmentation3d.loss.focal_loss import FocalLoss
    from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
    sd = oinit.randomize_affine(oinit.init_state_dict(arch, 1, cout, 0), 5)
    g = torch.Generator().manual_seed(21)
    crops = torch.randn((2, 1, 16, 16, 32), generator=g)
    masks = torch.randint(0, cout, (2, 1, 16, 16, 32), generator=g).float()
    # oracle: torch autograd through the functional restatement
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    probs = onet.forward_with_grad(params, crops)
    if lossname == 'dice':
        ref_loss = oloss.multi_dice_loss(probs, masks, [1.0] * cout)
    else:
        ref_loss = oloss.focal_loss(probs, masks, cout, alpha=[1.0] * cout, gamma=2)
    ref_loss.backward()
    # product
    net = _net(arch, cout, sd, 'fp32')
    out = net(crops.cuda())
    lf = MultiDiceLoss([1.0] * cout, cout, True) if lossname == 'dice' else FocalLoss(cout, alpha=[1.0] * cout, gamma=2, use_gpu=True)
    loss = lf(out, masks.cuda())
    loss.backward()
    assert abs(loss.item() - float(ref_loss)) <= 1e-5
    worst = 0.0
    for name, p in net.named_parameters():
        gr = params[name].grad
        assert p.grad is not None, name
        scale = float(gr.abs().max()) + 1e-12
        err = float((p.grad.cpu() - gr).abs().max()) / scale
        worst = max(worst, err)
        assert err <= 1e-2, (name, err, scale)   # Dice thresholds p > 1/C flip on near-tie voxels: not bit-stable across summation orders
    print(arch, 'worst relative grad error', worst)


def test_two_adam_steps_match_reference_golden():
    """core/seg_train.py:83,119-127 on the golden batch: losses and updated weights of the unmodified reference."""
    from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
    z = np.load(os.path.join(G, 'train_step.npz'))
    sd = oinit.init_state_dict('vnet', 1, 2, 0)
    net = _net('vnet', 2, sd, 'fp32')
    g = torch.Generator().manual_seed(21)
    crops = torch.randn((2, 1, 16, 16, 32), generator=g).cuda()
    masks = torch.randint(0, 2, (2, 1, 16, 16, 32), generator=g).float().cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999))
    lf = MultiDiceLoss([1.0, 1.0], 2, True)
    losses = []
    for _ in range(2):
        opt.zero_grad()
        loss = lf(net(crops), masks)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print('losses', losses, 'reference', z['vnet_dice_losses'])
    assert np.abs(np.array(losses) - z['vnet_dice_losses']).max() <= 2e-4
    got = net.state_dict()
    for k in ('in_block.conv.weight', 'down_64.down_conv.weight', 'up_32.up_conv.weight', 'out_block.conv2.weight',
              'out_block.gn1.weight', 'up_128.up_gn.bias'):
        ref = z['vnet_dice/' + k]
        d = np.abs(got[k].cpu().numpy() - ref)
        # Adam's first steps move every weight by ~lr; a wrong gradient sign shows up as a 2*lr = 2e-4 difference
        assert np.mean(d > 1.5e-4) <= 0.01, (k, float(d.max()), float(np.mean(d > 1.5e-4)))
    gi = net.in_block.conv.weight.grad.cpu().numpy()
    ref = z['vnet_dice_grad_in_block']
    assert np.abs(gi - ref).max() <= 5e-2 * np.abs(ref).max()
