"""Build-container check that the CPU baseline bench.py reports (the oracle PORT of core/seg_infer.segmentation_volume) runs
at the speed of the UNMODIFIED reference: both are timed here on the same synthetic volume, same weights, same partition.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/time_reference_cpu.py [size] [patch]

The reference is imported from /root/reference under the stand-ins of ref_shims.py (numpy-backed SimpleITK), so its
`sitk` calls are array copies exactly where the port makes its `faithful_copies`.  Prints both times and the ratio."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

sitk = ref_shims.install()
sys.dont_write_bytecode = True
size = int(sys.argv[1]) if len(sys.argv) > 1 else 192
patch = int(sys.argv[2]) if len(sys.argv) > 2 else 96

# ---- oracle port (what bench.py times) -------------------------------------------------------------------------------
sys.path.insert(0, ROOT)
from oracle import init as oinit                       # noqa: E402
from oracle import sliding_window as osw               # noqa: E402
sd = oinit.init_state_dict('vnet', 1, 2, 0)
g = torch.Generator().manual_seed(1)
vol = (torch.randn((size, size, size), generator=g) * 300).numpy().astype(np.float32)
norm = {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True}
t0 = time.time()
p_port, m_port, starts, _ = osw.segmentation_volume(sd, vol, [1.0, 1.0, 1.0], norm, 'SIZE', [patch] * 3, [patch] * 3, 16,
                                                    double_forward=True, faithful_copies=True)
t_port = time.time() - t0

# ---- the unmodified reference ------------------------------------------------------------------------------------------
sys.path.insert(0, '/root/reference')
from easydict import EasyDict as edict                                  # noqa: E402  (stand-in)
from segmentation3d.core import seg_infer as ref_infer                  # noqa: E402
from segmentation3d.network import vnet as ref_vnet                     # noqa: E402
from segmentation3d.utils.normalizer import FixedNormalizer             # noqa: E402
net = ref_vnet.SegmentationNet(1, 2)
net.load_state_dict(sd)
net.eval()
model = edict()
model.net = net
model.spacing, model.max_stride, model.interpolation = [1.0, 1.0, 1.0], 16, 'LINEAR'
model.in_channels, model.out_channels = 1, 2
model.crop_normalizers = [FixedNormalizer(0.0, 1000.0, True)]
cfg = edict()
cfg.partition_type, cfg.partition_size, cfg.partition_stride = 'SIZE', [patch] * 3, [patch] * 3
cfg.cpu_model_spacing_increase_ratio, cfg.cpu_partition_decrease_ratio = 1.0, 1.0
cfg.pick_largest_cc, cfg.remove_small_cc = False, 0
image = sitk.GetImageFromArray(vol)
t0 = time.time()
mean_probs, mask = ref_infer.segmentation_volume(model, cfg, image, None, None, False)
t_ref = time.time() - t0
p_ref = np.stack([sitk.GetArrayFromImage(p) for p in mean_probs], 0)
print('volume %d^3, %d patches of %d^3, %d threads' % (size, len(starts), patch, torch.get_num_threads()))
print('unmodified reference %.2f s, oracle port %.2f s, port / reference = %.3f' % (t_ref, t_port, t_port / t_ref))
print('max |dp| port vs reference %.3g, masks equal %s' % (float(np.abs(p_ref - p_port).max()),
                                                            bool(np.array_equal(sitk.GetArrayFromImage(mask).astype(np.int8), m_port))))

# ---- training step: unmodified reference modules vs the oracle port bench.py --task train times ---------------------------
from segmentation3d.loss.multi_dice_loss import MultiDiceLoss          # noqa: E402
from oracle import loss as oloss                                       # noqa: E402
from oracle import net as onet                                         # noqa: E402
g = torch.Generator().manual_seed(0)
crops = torch.randn((1, 1, patch, patch, patch), generator=g)
masks = torch.randint(0, 2, (1, 1, patch, patch, patch), generator=g).float()
net = ref_vnet.SegmentationNet(1, 2)
net.load_state_dict(sd)
net.train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999))
lf = MultiDiceLoss([0.5, 0.5], 2, False)
t0 = time.time()
opt.zero_grad()
loss_ref = lf(net(crops), masks)
loss_ref.backward()
opt.step()
t_ref = time.time() - t0
params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
opt = torch.optim.Adam(list(params.values()), lr=1e-4, betas=(0.9, 0.999))
t0 = time.time()
opt.zero_grad()
loss_port = oloss.multi_dice_loss(onet.forward_with_grad(params, crops), masks, [0.5, 0.5])
loss_port.backward()
opt.step()
t_port = time.time() - t0
print('training step on one %d^3 crop: unmodified reference %.2f s (loss %.7f), oracle port %.2f s (loss %.7f), port / reference = %.3f'
      % (patch, t_ref, float(loss_ref), t_port, float(loss_port), t_port / t_ref))
