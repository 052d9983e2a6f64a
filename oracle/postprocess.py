"""CPU restatement of the mask post-processing.  TEST INFRASTRUCTURE ONLY.

reference: segmentation3d/utils/image_tools.py:380-404 (pick_largest_connected_component) and :407-432
(remove_small_connected_component): per label, sitk.ConnectedComponentImageFilter with SetFullyConnected(True)
(26-connectivity) followed by sitk.RelabelComponent (components sorted by size; with a minimum object size the smaller ones are
dropped).  SimpleITK is absent here: scipy.ndimage.label with a full 3x3x3 structuring element computes the same components
(parity unpinned against SimpleITK itself; ties between equally large components go to the one met first in raster order).
"""
import numpy as np
from scipy import ndimage


def _filter(mask, labels, keep_threshold):
    out = np.zeros_like(mask)
    st = np.ones((3, 3, 3), dtype=bool)
    for lab in labels:
        cc, n = ndimage.label(mask == lab, structure=st)
        if n == 0:
            continue
        sizes = np.bincount(cc.ravel())[1:]
        if keep_threshold is None:
            keep = [int(np.argmax(sizes)) + 1]
        else:
            keep = [i + 1 for i, s in enumerate(sizes) if s >= keep_threshold]
        out[np.isin(cc, keep)] = lab
    return out


def pick_largest_connected_component(mask_zyx, labels):
    return _filter(np.asarray(mask_zyx), labels, None)


def remove_small_connected_component(mask_zyx, labels, threshold):
    return _filter(np.asarray(mask_zyx), labels, int(threshold))
