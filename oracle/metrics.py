"""Oracle: Dice ratio used for parity grading, and parity summary helpers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Reference: segmentation3d/utils/metrics.py:5-38 (cal_dsc).
"""
import numpy as np


def cal_dsc(gt_npy, seg_npy, label, threshold):
    """metrics.py:22-36."""
    gt, seg = (gt_npy == label), (seg_npy == label)
    area_gt, area_seg = np.sum(gt), np.sum(seg)
    if area_gt < threshold and area_seg < threshold:
        return 1.0, 'TN'
    if area_gt < threshold and area_seg >= threshold:
        return 0.0, 'FP'
    if area_gt >= threshold and area_seg < threshold:
        return 0.0, 'FN'
    inter = np.sum(gt & seg)
    return 2 * inter / (area_gt + area_seg), 'TP'


def parity_report(probs_ref, probs_new, threshold=1):
    """max|dp|, argmax agreement and per-class Dice of the argmax masks (BASELINE.json bars).
    Both inputs [C, ...] float32 numpy."""
    import torch
    pr = torch.from_numpy(np.ascontiguousarray(probs_ref))
    pn = torch.from_numpy(np.ascontiguousarray(probs_new))
    m_ref = pr.max(0)[1].numpy()
    m_new = pn.max(0)[1].numpy()
    rep = {
        'max_abs': float((pr - pn).abs().max()),
        'mean_abs': float((pr - pn).abs().mean()),
        'agree': float((m_ref == m_new).mean()),
        'dice': [float(cal_dsc(m_ref, m_new, c, threshold)[0]) for c in range(pr.shape[0])],
    }
    return rep
