"""torchrun --nproc-per-node N tools/check_patch_shard.py : one volume, patches dealt rank::N.  The reduce-scatter + slab
finalize + mask all-gather path must give the same mask as the all-reduce path and as a single-rank pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from segmentation3d.core.seg_infer import make_model, segmentation_volume_device
from segmentation3d.network import vnet

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda:%d' % local))
torch.manual_seed(0)
net = vnet.SegmentationNet(1, 2)
vnet.parameters_kaiming_init(net)
net.b200_mode = 'fp32x'
net = net.cuda().eval()
model = make_model(net, spacing=[1.0, 1.0, 1.0], normalizer={'type': 0, 'mean': 0.0, 'stddev': 1.0, 'clip': False})
cfg = {'partition_type': 'SIZE', 'partition_size': [48, 48, 48], 'partition_stride': [32, 32, 32]}
g = torch.Generator(device='cuda').manual_seed(7)
vol = torch.randn((64 * world // world * 2, 96, 112), generator=g, device='cuda')       # Z = 128: divisible by 2, 4, 8
vol = torch.nn.functional.avg_pool3d(vol[None, None], 3, 1, 1)[0, 0].contiguous() * 3.0
acc1, mask1 = segmentation_volume_device(model, cfg, vol, batch=4)
accp, maskp = segmentation_volume_device(model, cfg, vol, batch=4, shard=(rank, world), gather='probs')
accm, maskm = segmentation_volume_device(model, cfg, vol, batch=4, shard=(rank, world), gather='mask')
torch.cuda.synchronize()
zs = vol.shape[0] // world
d_probs = float((accp - acc1).abs().max())
d_slab = float((accm - acc1[:, rank * zs:(rank + 1) * zs]).abs().max())
agree_p = float((maskp == mask1).float().mean())
agree_m = float((maskm == mask1).float().mean())
print('rank %d/%d: all-reduce max|dp| %.3g, reduce-scatter slab max|dp| %.3g, mask agreement %.6f / %.6f'
      % (rank, world, d_probs, d_slab, agree_p, agree_m), flush=True)
assert d_probs <= 1e-5 and d_slab <= 1e-5 and agree_p >= 0.9999 and agree_m >= 0.9999
assert accm.shape == (2, zs) + tuple(vol.shape[1:]) and maskm.shape == vol.shape
# label exchange (gather='labels'): non-overlapping patches, local arg-max + max all-reduce of the int8 mask
from segmentation3d.core.seg_infer import _grid, deal_patches
cfg_t = {'partition_type': 'SIZE', 'partition_size': [48, 48, 48], 'partition_stride': [48, 48, 48]}
vol_t = vol[:96, :96, :96].contiguous()
st, en = _grid(model, cfg_t, [96, 96, 96], [1.0, 1.0, 1.0], None, None)
assert deal_patches(st, en, rank, world)[1] and len(st) == 8
_, mask_t1 = segmentation_volume_device(model, cfg_t, vol_t, batch=4)
_, mask_tl = segmentation_volume_device(model, cfg_t, vol_t, batch=4, shard=(rank, world), gather='labels')
torch.cuda.synchronize()
agree_l = float((mask_tl == mask_t1).float().mean())
print('rank %d/%d: label exchange mask agreement %.6f' % (rank, world, agree_l), flush=True)
assert agree_l >= 0.9999
# the same exchange on a volume whose ranks own different z ranges (3 x 2 x 2 patches of 48^3), device and HOST entry points:
# a rank accumulates, finishes and (host call) uploads only the z planes its patches touch
from segmentation3d.core.seg_infer import segmentation_volume_host, shard_plan
g2 = torch.Generator(device='cuda').manual_seed(9)
vol_z = torch.nn.functional.avg_pool3d(torch.randn((1, 1, 144, 96, 96), generator=g2, device='cuda'), 3, 1, 1)[0, 0].contiguous() * 3.0
_, _, mine, (z_lo, z_hi), disjoint = shard_plan(model, cfg_t, vol_z.shape, (rank, world))
assert disjoint
acc_z1, mask_z1 = segmentation_volume_device(model, cfg_t, vol_z, batch=4)
acc_zl, mask_zl = segmentation_volume_device(model, cfg_t, vol_z, batch=4, shard=(rank, world), gather='labels')
host_vol = torch.empty(vol_z.shape, dtype=torch.float32, pin_memory=True)
host_vol.copy_(vol_z)
host_mask = torch.empty(vol_z.shape, dtype=torch.int8, pin_memory=True)
_, host_mask = segmentation_volume_host(model, cfg_t, host_vol, host_mask, batch=4, shard=(rank, world), gather='labels')
torch.cuda.synchronize()
assert acc_zl.shape[1] == z_hi - z_lo and (world == 1 or z_hi - z_lo < 144), (acc_zl.shape, z_lo, z_hi)
own = torch.zeros(vol_z.shape, dtype=torch.bool, device='cuda')
for s0 in mine:
    own[s0[2]:s0[2] + 48, s0[1]:s0[1] + 48, s0[0]:s0[0] + 48] = True
d_own = float(((acc_zl - acc_z1[:, z_lo:z_hi]).abs() * own[z_lo:z_hi]).max())
agree_z = float((mask_zl == mask_z1).float().mean())
agree_h = float((host_mask.cuda() == mask_z1).float().mean())
print('rank %d/%d: z-range %d..%d of 144 (%d patches): own-patch max|dp| %.3g, label exchange agreement %.6f (device) %.6f (host call)'
      % (rank, world, z_lo, z_hi, len(mine), d_own, agree_z, agree_h), flush=True)
assert d_own <= 1e-5 and agree_z >= 0.9999 and agree_h >= 0.9999
# mask_root = 0: the merged mask is reduced to rank 0 alone, and only rank 0's host buffer receives it
host_mask0 = torch.full(vol_z.shape, -1, dtype=torch.int8).pin_memory()
_, out0 = segmentation_volume_host(model, cfg_t, host_vol, host_mask0, batch=4, shard=(rank, world), gather='labels', mask_root=0)
torch.cuda.synchronize()
if rank == 0:
    agree_0 = float((host_mask0.cuda() == mask_z1).float().mean())
    print('rank 0/%d: mask reduced to rank 0 only, agreement of its host mask %.6f' % (world, agree_0), flush=True)
    assert out0 is host_mask0 and agree_0 >= 0.9999
else:
    assert out0.is_cuda and int(host_mask0.max()) == -1          # the other ranks keep their partial device mask, no host copy
# ragged volume: the last box of every axis is clamped back into the volume and overlaps its neighbour (as in BASELINE
# configs[1]: 512 = 5 x 96 + 32); whole overlap components go to one rank, so the label exchange is still exact
vol_r = torch.nn.functional.avg_pool3d(torch.randn((1, 1, 160, 112, 128), generator=g2, device='cuda'), 3, 1, 1)[0, 0].contiguous() * 3.0
st_r, en_r, mine_r, zr, disjoint_r = shard_plan(model, cfg_t, vol_r.shape, (rank, world))
assert disjoint_r and len(st_r) == 4 * 3 * 3
_, mask_r1 = segmentation_volume_device(model, cfg_t, vol_r, batch=6)
_, mask_rl = segmentation_volume_device(model, cfg_t, vol_r, batch=6, shard=(rank, world), gather='labels')
torch.cuda.synchronize()
agree_r = float((mask_rl == mask_r1).float().mean())
print('rank %d/%d: ragged volume (36 patches, clamped last boxes overlap): %d patches here, z-range %s, label exchange agreement %.6f'
      % (rank, world, len(mine_r), zr, agree_r), flush=True)
assert agree_r >= 0.9999
dist.destroy_process_group()
