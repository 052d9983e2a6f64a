"""Two forwards of VNet on B random 96^3 patches (the second one is the one to capture under ncu):
    ncu --set full -k regex:<kernel> -s <launches of the first forward> ... python tools/profile_forward.py 8 fp16"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
from segmentation3d.network import vnet, vbnet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else 'fp16'
arch = sys.argv[3] if len(sys.argv) > 3 else 'vnet'
classes = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.manual_seed(0)
mod = vnet if arch == 'vnet' else vbnet
net = mod.SegmentationNet(1, classes)
mod.parameters_kaiming_init(net)
net.b200_mode = mode
net = net.cuda().eval()
x = torch.randn((B, 1, 96, 96, 96), device='cuda')
with torch.no_grad():
    for _ in range(2):
        y = net(x)
torch.cuda.synchronize()
print('ok', tuple(y.shape), float(y.sum()))
