"""torchrun --nproc-per-node 2 tools/check_overlap_allreduce.py : two data-parallel training steps with the gradient
all-reduce overlapped with the backward pass vs issued after it must leave identical parameters on every rank."""
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


def run(overlap, mode='fp32'):
    """averaged gradients of one data-parallel step (compared before Adam, which would amplify rounding noise on
    near-zero gradients to +-lr)"""
    os.environ['SEG3D_OVERLAP_ALLREDUCE'] = '1' if overlap else '0'
    torch.manual_seed(0)
    net = vnet.SegmentationNet(1, 2)
    vnet.parameters_kaiming_init(net)
    net.b200_mode = mode
    net = net.cuda().train()
    D.broadcast_params(net)
    lf = MultiDiceLoss([0.5, 0.5], 2, True)
    g = torch.Generator(device='cuda').manual_seed(100 + rank)
    crops = torch.randn((2, 1, 64, 64, 64), generator=g, device='cuda')
    masks = torch.randint(0, 2, (2, 1, 64, 64, 64), generator=g, device='cuda').float()
    params = list(net.parameters())
    loss = lf(net(crops), masks)
    loss.backward()
    reduced = net._plan.grads_reduced_in_backward
    if not reduced:
        D.allreduce_mean_grads(params)
    torch.cuda.synchronize()
    return [p.grad.detach().clone() for p in params], float(loss), reduced


# The exchange is checked in the strict fp32 mode, where a step is reproducible to fp32-atomic rounding (~1e-6).  In bf16 two
# runs of the SAME schedule already differ by up to ~1e-2 in the deep layers in about half of the processes: the double-precision
# atomics of the loss / tail reductions land in a different order, one stored bf16 gradient flips by an ulp, and the deep layers
# of a random-init net amplify that (tools/diag_overlap.py shows the first differing tensor is the tail's
# gradient, with the forward sums bit-identical) - so bf16 gets the band of that noise, not the 1e-3 of a race check.
ga, la, fa = run(True)
gb, lb, fb = run(False)
assert fa and not fb, (fa, fb)
worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-20)) for a, b in zip(ga, gb))
ha, _, _ = run(True, 'bf16')
hb, _, _ = run(False, 'bf16')
worst_bf16 = max(float((a - b).abs().max() / (b.abs().max() + 1e-20)) for a, b in zip(ha, hb))
print('rank %d: bf16 step, overlapped vs sequential: worst per-tensor relative gradient difference %.3g (run-to-run band of bf16)'
      % (rank, worst_bf16), flush=True)
assert worst_bf16 <= 5e-2
if worst > 1e-4:        # name the tensors: a race between an all-reduce and a kernel still writing its range shows up here
    _net = vnet.SegmentationNet(1, 2)
    names = [n for n, _ in _net.named_parameters()]
    rows = sorted(((float((a - b).abs().max() / (b.abs().max() + 1e-20)), n) for n, a, b in zip(names, ga, gb)), reverse=True)
    print('rank %d: largest differences: %s' % (rank, ', '.join('%s %.2g' % (n, r) for r, n in rows[:6])), flush=True)
flat = torch.cat([g.reshape(-1) for g in ga])
other = flat.clone()
dist.broadcast(other, 0)
same_across_ranks = bool(torch.equal(other, flat))
print('rank %d: overlapped vs sequential all-reduce, worst per-tensor relative gradient difference %.3g, losses %.6f / %.6f, '
      'ranks identical: %s' % (rank, worst, la, lb, same_across_ranks), flush=True)
# wgrad accumulates with fp32 atomics (order varies run to run), so the two runs agree to rounding, not bitwise
assert worst <= 1e-4 and same_across_ranks


def dp_vs_global_batch():
    """data-parallel step (per-rank batch 2, gradients averaged over ranks) against ONE process running the global batch, in the
    strict fp32 mode: the reference's single-process DataParallel step takes the loss mean over the global batch
    (core/seg_train.py:77,119-127), and the mean of equal-sized per-rank means is that mean."""
    world = dist.get_world_size()
    os.environ['SEG3D_OVERLAP_ALLREDUCE'] = '1'

    def data(r):
        g = torch.Generator(device='cuda').manual_seed(200 + r)
        return (torch.randn((2, 1, 32, 32, 32), generator=g, device='cuda'),
                torch.randint(0, 2, (2, 1, 32, 32, 32), generator=g, device='cuda').float())

    torch.manual_seed(0)
    net = vnet.SegmentationNet(1, 2)
    vnet.parameters_kaiming_init(net)
    net.b200_mode = 'fp32'
    net = net.cuda().train()
    lf = MultiDiceLoss([0.5, 0.5], 2, True)
    crops, masks = data(rank)
    loss = lf(net(crops), masks)
    loss.backward()
    params = list(net.parameters())
    if not net._plan.grads_reduced_in_backward:
        D.allreduce_mean_grads(params)
    torch.cuda.synchronize()
    return [p.grad.detach().clone() for p in params], float(loss), world, data


g_dp, l_dp, world, data = dp_vs_global_batch()
# the global-batch arm runs outside the process group (a world of one) so that nothing is averaged across ranks
dist.barrier()
dist.destroy_process_group()
torch.manual_seed(0)
net = vnet.SegmentationNet(1, 2)
vnet.parameters_kaiming_init(net)
net.b200_mode = 'fp32'
net = net.cuda().train()
lf = MultiDiceLoss([0.5, 0.5], 2, True)
allc = torch.cat([data(r)[0] for r in range(world)], 0)
allm = torch.cat([data(r)[1] for r in range(world)], 0)
loss = lf(net(allc), allm)
loss.backward()
torch.cuda.synchronize()
worst_dp = max(float((a - p.grad).abs().max() / (p.grad.abs().max() + 1e-20)) for a, p in zip(g_dp, net.parameters()))
print('rank %d: data-parallel (world %d, batch 2/rank) vs one process on the global batch of %d, fp32: worst per-tensor relative '
      'gradient difference %.3g, global loss %.6f' % (rank, world, 2 * world, worst_dp, float(loss)), flush=True)
assert worst_dp <= 2e-3
