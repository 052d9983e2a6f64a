"""Dice ratio (drop-in for reference utils/metrics.py:5-38)."""
import numpy as np


def _as_array(x):
    if hasattr(x, 'to_numpy'):
        return x.to_numpy()
    try:
        import SimpleITK as sitk
        if isinstance(x, sitk.Image):
            return sitk.GetArrayFromImage(x)
    except ImportError:
        pass
    return np.asarray(x)


def cal_dsc(gt_npy, seg_npy, label, threshold):
    """(dsc, seg_type) with seg_type in TN / FP / FN / TP; a structure smaller than `threshold`
    voxels counts as absent."""
    gt, seg = _as_array(gt_npy) == label, _as_array(seg_npy) == label
    n_gt, n_seg = int(gt.sum()), int(seg.sum())
    if n_gt < threshold:
        return (1.0, 'TN') if n_seg < threshold else (0.0, 'FP')
    if n_seg < threshold:
        return 0.0, 'FN'
    return 2 * int((gt & seg).sum()) / (n_gt + n_seg), 'TP'
