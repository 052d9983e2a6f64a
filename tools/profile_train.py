"""Per-entry-point time breakdown of one training step (synchronising event pair around every C-ABI call)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
from segmentation3d._b200 import lib
from segmentation3d.core.seg_train import make_optimizer, train_step
from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
from segmentation3d.network import vnet

mode = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
net = vnet.SegmentationNet(1, 2)
vnet.parameters_kaiming_init(net)
net.b200_mode = mode
net = net.cuda().train()
opt = make_optimizer(net, 1e-4)
lf = MultiDiceLoss([0.5, 0.5], 2, True)
crops = torch.randn((B, 1, 96, 96, 96), device='cuda')
masks = torch.randint(0, 2, (B, 1, 96, 96, 96), device='cuda').float()
for _ in range(2):
    train_step(net, opt, lf, crops, masks)
torch.cuda.synchronize()
times = collections.defaultdict(float)
counts = collections.defaultdict(int)
orig = lib.call


def timed(name, *args):
    key = name
    if name == 'seg3d_conv3d_fwd':
        key = name + ('(tc)' if args[2] == lib.IMPL_TCGEN05 else '(simt)') + ('[stats]' if args[15] else '[dgrad]')
    if name == 'seg3d_conv3d_wgrad':
        key = 'wgrad mode%d %dx%d D=%d' % (args[0], args[4], args[7], args[10])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig(name, *args)
    b.record()
    torch.cuda.synchronize()
    times[key] += a.elapsed_time(b)
    counts[key] += 1


lib.call = timed
import segmentation3d._b200.plan as P, segmentation3d._b200.autograd as A, segmentation3d.loss._kernels as K
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
train_step(net, opt, lf, crops, masks)
t1.record()
torch.cuda.synchronize()
tot = sum(times.values())
print('mode %s B=%d: step (with per-call syncs) %.1f ms, kernels %.1f ms' % (mode, B, t0.elapsed_time(t1), tot))
for k, v in sorted(times.items(), key=lambda kv: -kv[1]):
    print('%-40s n=%3d %8.2f ms %5.1f%%' % (k, counts[k], v, 100 * v / tot))
