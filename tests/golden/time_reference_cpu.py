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
