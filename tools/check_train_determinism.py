"""python tools/check_train_determinism.py : the same bf16 training step twice on one GPU; per-tensor relative gradient differences
(weight gradients accumulate with fp32 atomics, so the runs agree to rounding, not bitwise)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
from segmentation3d.network import vnet


def run():
    torch.manual_seed(0)
    net = vnet.SegmentationNet(1, 2)
    vnet.parameters_kaiming_init(net)
    net.b200_mode = 'bf16'
    net = net.cuda().train()
    lf = MultiDiceLoss([0.5, 0.5], 2, True)
    g = torch.Generator(device='cuda').manual_seed(100)
    crops = torch.randn((2, 1, 64, 64, 64), generator=g, device='cuda')
    masks = torch.randint(0, 2, (2, 1, 64, 64, 64), generator=g, device='cuda').float()
    loss = lf(net(crops), masks)
    loss.backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in net.named_parameters()}, float(loss)


ga, la = run()
gb, lb = run()
rows = sorted(((float((ga[n] - gb[n]).abs().max() / (gb[n].abs().max() + 1e-20)), n, float(gb[n].abs().max())) for n in ga), reverse=True)
print('losses %.7f %.7f' % (la, lb))
for r, n, m in rows[:8]:
    print('%-40s rel diff %.3g   max|g| %.3g' % (n, r, m))
