"""A few VNet training steps (B patches of 96^3, Dice, Adam) for ncu launch lists:
    ncu --metrics gpu__time_duration.sum -s <launches of the warm-up steps> python tools/train_one_step.py bf16 8 3"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
from segmentation3d.core.seg_train import make_optimizer, train_step
from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
from segmentation3d.network import vnet

mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
net = vnet.SegmentationNet(1, 2)
vnet.parameters_kaiming_init(net)
net.b200_mode = mode
net = net.cuda().train()
opt = make_optimizer(net, 1e-4)
lf = MultiDiceLoss([0.5, 0.5], 2, True)
crops = torch.randn((B, 1, 96, 96, 96), device='cuda')
masks = torch.randint(0, 2, (B, 1, 96, 96, 96), device='cuda').float()
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push('last_step')
    loss = train_step(net, opt, lf, crops, masks)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print('loss', float(loss))
