"""torchrun --nproc-per-node 2 tools/diag_overlap.py : which part of a bf16 data-parallel step is not reproducible between the
overlapped and the sequential gradient all-reduce - forward GroupNorm sums, or gradients only?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from segmentation3d._b200 import dist as D
from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
from segmentation3d.network import vnet

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))


def run(overlap):
    os.environ['SEG3D_OVERLAP_ALLREDUCE'] = '1' if overlap else '0'
    torch.manual_seed(0)
    net = vnet.SegmentationNet(1, 2)
    vnet.parameters_kaiming_init(net)
    net.b200_mode = 'bf16'
    net = net.cuda().train()
    D.broadcast_params(net)
    lf = MultiDiceLoss([0.5, 0.5], 2, True)
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    crops = torch.randn((2, 1, 64, 64, 64), generator=g, device='cuda')
    masks = torch.randint(0, 2, (2, 1, 64, 64, 64), generator=g, device='cuda').float()
    params = list(net.parameters())
    probs = net(crops)
    plan = net._plan
    ws = [w for w, _ in plan._plans.values()][-1]
    stats = ws['stats'].clone()
    loss = lf(probs, masks)
    loss.backward()
    if not plan.grads_reduced_in_backward:
        D.allreduce_mean_grads(params)
    torch.cuda.synchronize()
    names = [n for n, _ in net.named_parameters()]
    bw = ws['bwd']
    path = [('tail.gy', bw.gy_tail.buf), ('tail.gd', bw.gd_tail.buf)]
    for u in reversed(bw.units):                       # backward order
        k = u['conv']
        path.append((k + '.gy', bw.gy[k].buf))
        if k in bw.dres:
            path.append((k + '.dres', bw.dres[k].buf))
        if k in bw.gd:
            path.append((k + '.gd', bw.gd[k].buf))
    sig = [(n, float(t.float().abs().double().sum()), float(t.float().double().sum())) for n, t in path]
    return {n: p.grad.detach().clone() for n, p in zip(names, params)}, stats, plan.gn_names, float(loss), sig


runs = [run(o) for o in (True, False, True, False, False)]
tags = ['ovl1', 'seq1', 'ovl2', 'seq2', 'seq3']
ref_sig = runs[0][4]
for r in range(1, len(runs)):
    first = None
    for (n, a0, b0), (_, a1, b1) in zip(ref_sig, runs[r][4]):
        if abs(a0 - a1) > 1e-12 * max(abs(a0), 1e-30) or abs(b0 - b1) > 1e-12 * max(abs(a0), 1e-30):
            first = (n, a0, a1)
            break
    print('rank %d data path of %s vs ovl1: first differing tensor in backward order: %s' % (rank, tags[r], first), flush=True)
for i in range(len(runs)):
    for j in range(i + 1, len(runs)):
        ga, sa, names, la = runs[i][:4]
        gb, sb, _, lb = runs[j][:4]
        gw = max(float((ga[n] - gb[n]).abs().max() / (gb[n].abs().max() + 1e-20)) for n in ga)
        sd = ((sa - sb).abs() / (sb.abs() + 1e-30)).amax(dim=(1, 2))
        k = int(sd.argmax())
        first = next((names[t] for t in range(len(names)) if float(sd[t]) > 1e-9), '-')
        print('rank %d %s vs %s: worst grad rel diff %.3g | forward GN sums: worst rel diff %.3g at %s, some layer above 1e-9: %s' %
              (rank, tags[i], tags[j], gw, float(sd[k]), names[k], first), flush=True)
