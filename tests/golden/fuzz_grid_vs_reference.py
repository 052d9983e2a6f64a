"""Build-container fuzz: the oracle's partition_grid against the UNMODIFIED reference image_partition_by_fixed_size
(utils/image_tools.py:163-218) on random volumes / spacings / bounding boxes / partition sizes and strides.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/fuzz_grid_vs_reference.py [n]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

sitk = ref_shims.install()
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
from oracle import sliding_window as osw                     # noqa: E402
sys.path.insert(0, '/root/reference')
from segmentation3d.utils import image_tools as ref_it       # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rng = np.random.default_rng(7)
same = raised = 0
for _ in range(n):
    size = [int(16 * rng.integers(1, 14)) for _ in range(3)]
    spacing = [float(rng.choice([0.3, 0.4, 0.5, 0.8, 1.0, 1.25, 2.0, 3.0])) for _ in range(3)]
    if rng.random() < 0.4:
        bs, be = [0, 0, 0], list(size)
    else:
        bs = [int(rng.integers(0, size[a] - 1)) for a in range(3)]
        be = [int(rng.integers(bs[a] + 1, size[a] + 1)) for a in range(3)]
    psize = [float(rng.choice([16, 24, 32, 48, 64, 96, 51.2, 89.6, 128])) for _ in range(3)]
    pstride = [float(max(2.0, psize[a] * rng.choice([0.2, 0.25, 0.5, 0.75, 1.0, 1.5]))) for a in range(3)]
    image = sitk.Image(size, sitk.sitkFloat32)
    image.SetSpacing(spacing)
    b1, e1, b2, e2 = list(bs), list(be), list(bs), list(be)
    try:
        ref = ref_it.image_partition_by_fixed_size(image, b1, e1, list(psize), list(pstride), 16)
    except AssertionError:
        try:
            osw.partition_grid(size, spacing, b2, e2, list(psize), list(pstride), 16)
            raise SystemExit('reference asserted, oracle did not: %r' % ((size, spacing, bs, be, psize, pstride),))
        except AssertionError:
            raised += 1
            continue
    got = osw.partition_grid(size, spacing, b2, e2, list(psize), list(pstride), 16)
    ok = ([list(map(int, s)) for s in got[0]] == [list(map(int, s)) for s in ref[0]] and
          [list(map(int, s)) for s in got[1]] == [list(map(int, s)) for s in ref[1]] and
          [int(v) for v in b1] == [int(v) for v in b2] and [int(v) for v in e1] == [int(v) for v in e2])
    if not ok:
        raise SystemExit('MISMATCH %r' % ((size, spacing, bs, be, psize, pstride),))
    same += 1
print('%d random configurations: %d identical grids (boxes, order, in-place bbox update), %d where both assert' % (n, same, raised))
