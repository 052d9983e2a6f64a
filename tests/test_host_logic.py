"""CPU tests of the host-side logic added around the hot path (no GPU, no CUDA calls)."""
import os

import numpy as np
import pytest
import torch

from oracle import postprocess as opp
from oracle import resample as orz
from segmentation3d._b200.sliding import axis_counts, batch_ranges
from segmentation3d.utils.image3d import Image3d
from segmentation3d.utils.image_tools import is_identity_resample, resample_size


def test_batch_ranges_are_balanced_and_cover_everything():
    for n in list(range(0, 50)) + [180, 181, 799, 800]:
        for batch in (1, 3, 6, 20):
            r = batch_ranges(n, batch)
            assert sum(c for _, c in r) == n
            assert all(0 < c <= batch for _, c in r)
            assert [s for s, _ in r] == list(np.cumsum([0] + [c for _, c in r])[:-1])
            if r:
                assert max(c for _, c in r) - min(c for _, c in r) <= max(c for _, c in r) - 1   # no degenerate tails ...
                assert len(r) == -(-n // batch)                                                     # ... and no extra forwards
    assert batch_ranges(23, 20) == [(0, 12), (12, 11)]
    assert batch_ranges(180, 20) == [(i * 20, 20) for i in range(9)]


def test_resample_size_matches_oracle_and_reference_rule():
    rng = np.random.default_rng(0)
    for _ in range(200):
        size = [int(v) for v in rng.integers(8, 600, 3)]
        sp_in = [float(v) for v in rng.uniform(0.2, 3.0, 3)]
        sp_out = [float(v) for v in rng.uniform(0.2, 3.0, 3)]
        assert resample_size(size, sp_in, sp_out, 16) == orz.out_size(size, sp_in, sp_out, 16)
    assert resample_size([512, 512, 400], [0.4, 0.4, 0.4], [0.4, 0.4, 0.4], 16) == [512, 512, 400]
    im = Image3d(np.zeros((400, 512, 512), np.float32), (1.0, 1.0, 1.0))
    assert is_identity_resample(im, [1.0, 1.0, 1.0], 16)
    assert not is_identity_resample(im, [1.0, 1.0, 2.0], 16)
    assert not is_identity_resample(Image3d(np.zeros((40, 50, 64), np.float32)), [1.0, 1.0, 1.0], 16)   # 50 % 16 != 0


def test_image3d_physical_point_round_trip():
    im = Image3d(np.zeros((10, 12, 14), np.float32), (0.5, 0.75, 2.0), (10.0, -5.0, 3.0))
    for idx in ([0, 0, 0], [3, 7, 9], [13, 11, 9]):
        pt = im.TransformContinuousIndexToPhysicalPoint([float(v) for v in idx])
        assert tuple(im.TransformPhysicalPointToIndex(pt)) == tuple(idx)
    assert np.allclose(im.TransformContinuousIndexToPhysicalPoint([2.0, 4.0, 1.0]), [11.0, -2.0, 5.0])


def test_axis_counts_factorise_the_overlap_count():
    from oracle import sliding_window as osw
    size = [80, 64, 48]
    starts, ends = osw.partition_grid(size, [1, 1, 1], [0, 0, 0], list(size), [32, 32, 32], [16, 24, 32], 16)
    cx, cy, cz = axis_counts(size, starts, ends)
    cnt = osw.overlap_count_axes(size, starts, ends)
    assert np.array_equal(cnt, (cz[:, None, None] * cy[None, :, None] * cx[None, None, :]).astype(np.float32))


def test_oracle_connected_component_restatement():
    m = np.zeros((6, 8, 10), np.int8)
    m[0, 0, 0:3] = 1                 # 3 voxels
    m[2:5, 2:5, 2:5] = 1             # 27 voxels, diagonal neighbour of nothing else
    m[5, 5, 5] = 1                   # touches the cube only by a corner: 26-connectivity joins them (28 voxels)
    m[0, 7, 9] = 2
    out = opp.pick_largest_connected_component(m, [1, 2])
    assert int((out == 1).sum()) == 28 and out[5, 5, 5] == 1 and out[0, 0, 0] == 0 and out[0, 7, 9] == 2
    out = opp.remove_small_connected_component(m, [1, 2], 3)
    assert int((out == 1).sum()) == 31 and out[0, 7, 9] == 0
    assert np.array_equal(opp.remove_small_connected_component(m, [1, 2], 1), m)


def test_make_optimizer_matches_reference_settings():
    from segmentation3d.core.seg_train import make_optimizer
    net = torch.nn.Linear(3, 2)
    opt = make_optimizer(net, 1e-4, (0.9, 0.999))
    g = opt.param_groups[0]
    assert isinstance(opt, torch.optim.Adam) and g['lr'] == 1e-4 and tuple(g['betas']) == (0.9, 0.999)
    assert not g.get('fused')        # fused only for CUDA parameters; the update rule is the same either way


def test_case_list_is_dealt_round_robin_over_ranks(monkeypatch):
    """BASELINE configs[4]: a txt case list is sharded by case over the ranks of a torchrun launch, no collective."""
    from segmentation3d.core.seg_infer import launch_rank, shard_case_list
    names, paths = ['c%d' % i for i in range(11)], ['/p/%d.mha' % i for i in range(11)]
    assert shard_case_list(names, paths, 0, 1) == (names, paths)
    seen = []
    for r in range(4):
        n, p = shard_case_list(names, paths, r, 4)
        assert n == names[r::4] and p == paths[r::4]
        assert len(n) in (2, 3)                                     # balanced to within one case
        seen += n
    assert sorted(seen) == sorted(names)                            # every case exactly once
    assert shard_case_list(names[:2], paths[:2], 3, 4) == ([], [])  # more ranks than cases: idle ranks get nothing
    with pytest.raises(ValueError):
        shard_case_list(names, paths, 4, 4)
    for k in ('WORLD_SIZE', 'RANK', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    assert launch_rank() == (0, 1, None)
    monkeypatch.setenv('WORLD_SIZE', '8'); monkeypatch.setenv('RANK', '5'); monkeypatch.setenv('LOCAL_RANK', '5')
    assert launch_rank() == (5, 8, 5)


def test_seg_eval_batch_dice_table(tmp_path):
    """core/seg_eval.py::cal_dsc_batch: per-case Dice / type per label plus mean and std rows, scored like the oracle's
    cal_dsc restatement (utils/metrics.py:22-36)."""
    import pandas as pd
    from oracle.metrics import cal_dsc as oracle_dsc
    from segmentation3d.core.seg_eval import cal_dsc_batch
    from segmentation3d.utils.image3d import write_image
    rng = np.random.default_rng(3)
    gts, segs, expect = [], [], []
    for i in range(3):
        gt = (rng.random((12, 16, 20)) > 0.6).astype(np.int8) * (1 + (rng.random((12, 16, 20)) > 0.5)).astype(np.int8)
        seg = gt.copy()
        seg[rng.random(gt.shape) > 0.9] = 0
        if i == 2:
            seg[seg == 2] = 0                                      # label 2 missed entirely -> FN
        os.makedirs(tmp_path / ('c%d' % i))
        g, s = str(tmp_path / ('c%d' % i) / 'gt.mha'), str(tmp_path / ('c%d' % i) / 'seg.mha')
        write_image(Image3d(gt), g, True)
        write_image(Image3d(seg), s, True)
        gts.append(g); segs.append(s)
        expect.append([oracle_dsc(gt, seg, l, 10) for l in (1, 2, 3)])
    df = cal_dsc_batch(gts, segs, [1, 2, 3], 10, str(tmp_path / 'res.csv'))
    back = pd.read_csv(str(tmp_path / 'res.csv'), index_col=0)
    assert list(back.columns) == ['filename', 'label1_score', 'label1_type', 'label2_score', 'label2_type', 'label3_score', 'label3_type']
    assert list(back['filename']) == ['gt.mha'] * 3 + ['mean', 'std']
    for i in range(3):
        for j, l in enumerate((1, 2, 3)):
            assert abs(back['label%d_score' % l].iloc[i] - expect[i][j][0]) < 1e-12
            assert back['label%d_type' % l].iloc[i] == expect[i][j][1]
    assert back['label2_type'].iloc[2] == 'FN' and back['label3_type'].iloc[0] == 'TN'
    assert abs(back['label1_score'].iloc[3] - np.mean([e[0][0] for e in expect])) < 1e-12
    assert abs(back['label1_score'].iloc[4] - np.std([e[0][0] for e in expect], ddof=1)) < 1e-12
    assert len(df) == 5


def test_bounding_box_grid_and_counts_match_oracle():
    """Host side of the cascade call (core/seg_infer.py:292-307): the product's patch grid for a bounding box equals the
    oracle's (pinned to the reference by tests/golden/cascade.npz), and the separable per-axis overlap count equals the
    rasterised one, 0 outside the visited box."""
    from oracle import sliding_window as osw
    from segmentation3d.core.seg_infer import _grid
    size, spacing = [64, 48, 64], [1.0, 1.0, 1.0]
    cfg = {'partition_type': 'SIZE', 'partition_size': [32, 32, 32], 'partition_stride': [16, 16, 16]}
    for bs, be in (([9, 5, 14], [49, 40, 50]), ([0, 0, 0], [64, 48, 64]), ([40, 30, 50], [64, 48, 64]), ([3, 3, 3], [20, 19, 18])):
        starts, ends = _grid({'max_stride': 16}, cfg, size, spacing, list(bs), list(be))
        o_starts, o_ends = osw.partition_grid(size, spacing, list(bs), list(be), [32, 32, 32], [16, 16, 16], 16)
        assert [list(map(int, s)) for s in starts] == [list(map(int, s)) for s in o_starts]
        assert [list(map(int, e)) for e in ends] == [list(map(int, e)) for e in o_ends]
        cx, cy, cz = axis_counts(size, starts, ends)
        sep = cz[:, None, None].astype(np.float32) * cy[None, :, None] * cx[None, None, :]
        assert np.array_equal(sep, osw.overlap_count_axes(size, o_starts, o_ends))
