"""Inference configuration template copied into every model folder by `train`
(same keys as reference config/infer_config.py:3-113)."""
from easydict import EasyDict as edict

__C = edict()
cfg = __C

__C.general = {}
# 'coarse' | 'fine' : run only that model;  'DISABLE' : coarse model, bounding box of its mask, then fine model
__C.general.single_scale = 'DISABLE'


def _scale(name, partition_type, size_mm):
    s = edict()
    s.model_name = name                         # sub-folder holding checkpoints/chk_<epoch>/params.pth
    s.pick_largest_cc = True                    # keep the largest 26-connected component per label
    s.remove_small_cc = 0                       # > 0: drop components smaller than this many voxels
    s.partition_type = partition_type           # 'SIZE' sliding window | 'DISABLE' whole volume in one forward
    s.partition_size = [size_mm] * 3            # mm
    s.partition_stride = [size_mm] * 3          # mm; smaller than the size => overlapping windows are averaged
    s.cpu_model_spacing_increase_ratio = 1.0    # only read when gpu_id == 0 selects the reference's "cpu knobs"
    s.cpu_partition_decrease_ratio = 1.0
    return s


__C.coarse = _scale('coarse', 'DISABLE', 51.2)
__C.fine = _scale('fine', 'SIZE', 89.6)
