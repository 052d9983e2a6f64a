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


def _dataset_cases():
    """same synthetic cases as tests/golden/make_golden.py::dataset_cases"""
    rng = np.random.RandomState(77)
    cases = []
    for k, (size, spacing, origin) in enumerate((((40, 36, 30), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)),
                                                 ((48, 40, 20), (0.8, 0.8, 1.5), (-10.0, 5.0, 30.0)),
                                                 ((30, 44, 38), (1.2, 0.9, 1.0), (3.5, -7.25, 12.0)))):
        im = rng.standard_normal((size[2], size[1], size[0])).astype(np.float32)
        lab = rng.randint(0, 3 if k != 2 else 2, size=(size[2], size[1], size[0])).astype(np.float32)
        lab[rng.random_sample(lab.shape) < 0.7] = 0
        cases.append((im, lab, spacing, origin))
    return cases


def test_dataset_requests_the_reference_crops(tmp_path):
    """dataloader/dataset.py:140-209: under the same numpy seed the drop-in dataset asks for exactly the crops (origin,
    spacing, size, interpolator), frames and case names the unmodified reference asked for
    (tests/golden/dataset_sampling.json) - all four sampling methods, images thinner than the crop, absent labels."""
    import json
    from segmentation3d.dataloader.dataset import SegmentationDataset
    from segmentation3d.utils import image_tools
    from segmentation3d.utils.image3d import write_image
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'dataset_sampling.json')))
    lines = ['3']
    for k, (im, lab, spacing, origin) in enumerate(_dataset_cases()):
        d = tmp_path / ('case%d' % k)
        os.makedirs(d)
        write_image(Image3d(im, spacing, origin), str(d / 'im.mha'), False)
        write_image(Image3d(lab, spacing, origin), str(d / 'seg.mha'), True)
        lines += [str(d / 'im.mha'), str(d / 'seg.mha')]
    with open(str(tmp_path / 'train.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    calls = []
    real_geometry = image_tools.crop_geometry

    def recording_geometry(center, size, spacing):
        out = real_geometry(center, size, spacing)
        calls.append(out)
        return out
    image_tools.crop_geometry = recording_geometry
    try:
        for method, items in gold.items():
            ds = SegmentationDataset(str(tmp_path / 'train.txt'), num_classes=3, spacing=[1.0, 1.0, 1.2], crop_size=[32, 32, 32],
                                     sampling_method=method, random_translation=[5, 4, 3], random_scale=[0.9, 1.1],
                                     interpolation='LINEAR', crop_normalizers=[None])
            np.random.seed(1234)
            for item in items:
                del calls[:]
                im_t, seg_t, frame, name = ds[item['index']]
                assert tuple(im_t.shape) == (1, 32, 32, 32) and tuple(seg_t.shape) == (1, 32, 32, 32)
                assert im_t.dtype == torch.float32 and seg_t.dtype == torch.float32
                assert name == item['name'] and len(calls) == 2
                for (origin, size, spacing), ref in zip(calls, item['calls']):
                    assert size == ref['size']
                    assert np.allclose(origin, ref['origin'], rtol=0, atol=1e-9), (method, item['index'])
                    assert np.allclose(spacing, ref['spacing'], rtol=0, atol=1e-12)
                assert np.allclose(frame, np.array(item['frame'], dtype=np.float32), rtol=0, atol=1e-5)
                assert set(np.unique(seg_t.numpy())) <= {0.0, 1.0, 2.0}          # nearest-neighbour mask crop keeps labels
    finally:
        image_tools.crop_geometry = real_geometry


def test_crop_image_restates_itk_resample():
    """utils/image_tools.py:111-146 (sitk.Resample, identity transform, default pixel 0): a crop on the source lattice
    is a copy with zeros outside the volume, a rescaled crop that starts at the volume's first voxel equals the
    resampling restatement used for the inference side (oracle/resample.py), 'NN' keeps labels, and a rotated
    direction matrix is honoured."""
    from segmentation3d.utils.image_tools import crop_image
    rng = np.random.default_rng(5)
    src = rng.standard_normal((20, 24, 28)).astype(np.float32)                  # z, y, x
    spacing, origin = (0.8, 1.25, 2.0), (-4.0, 10.0, 3.0)
    img = Image3d(src, spacing, origin)
    # (a) on-lattice crop of 16^3 voxels whose first voxel is source voxel (x=20, y=-3, z=8): copy + zero padding
    first = np.array([20, -3, 8])
    center = np.array(origin) + (first + 16 / 2.0 - 0.5) * np.array(spacing)
    for interp in ('LINEAR', 'NN'):
        out = crop_image(img, center, [16, 16, 16], spacing, interp)
        ref = np.zeros((16, 16, 16), np.float32)
        ref[:12, 3:, :8] = src[8:20, 0:13, 20:28]
        assert np.abs(out.to_numpy() - ref).max() <= 1e-6, interp
        assert np.allclose(out.GetOrigin(), np.array(origin) + first * np.array(spacing)) and np.allclose(out.GetSpacing(), spacing)
    # (b) rescaled crop anchored at the first voxel == oracle resample_grid (same origin, ratio = crop spacing / spacing)
    csp = (1.0, 1.0, 1.5)
    size = [16, 24, 20]
    center = [origin[a] + size[a] * csp[a] / 2.0 - csp[a] / 2.0 for a in range(3)]
    for interp in ('LINEAR', 'NN'):
        out = crop_image(img, center, size, csp, interp).to_numpy()
        ref = orz.resample_grid(src, spacing, size, csp, interp, 0.0)
        assert np.abs(out - ref).max() <= 1e-5, interp
    # (c) just below the first voxel centre (continuous index in [-0.5, 0)) linear interpolation clamps to the edge value
    out = crop_image(img, [origin[0] - 0.3 * spacing[0], origin[1], origin[2]], [1, 1, 1], spacing, 'LINEAR').to_numpy()
    assert abs(float(out[0, 0, 0]) - float(src[0, 0, 0])) <= 1e-6
    out = crop_image(img, [origin[0] - 0.6 * spacing[0], origin[1], origin[2]], [1, 1, 1], spacing, 'LINEAR').to_numpy()
    assert float(out[0, 0, 0]) == 0.0
    # (d) direction matrix: x and y axes swapped -> the crop walks the volume's own axes
    d = (0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0)
    rot = Image3d(src, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), d)
    p = rot.TransformContinuousIndexToPhysicalPoint([5.0, 7.0, 9.0])               # world position of voxel (x=5, y=7, z=9)
    out = crop_image(rot, p, [1, 1, 1], (1.0, 1.0, 1.0), 'NN')
    assert float(out.to_numpy()[0, 0, 0]) == float(src[9, 7, 5])


def test_dataset_batches_through_dataloader(tmp_path):
    """core/seg_train.py:84-87,119: EpochConcateSampler + DataLoader collate the items into
    (crops [B,1,D,H,W] f32, masks [B,1,D,H,W] f32, frames [B,15], names) - with worker processes too."""
    from torch.utils.data import DataLoader
    from segmentation3d.dataloader.dataset import SegmentationDataset
    from segmentation3d.dataloader.sampler import EpochConcateSampler
    from segmentation3d.utils.image3d import write_image
    from segmentation3d.utils.normalizer import FixedNormalizer
    lines = ['3']
    for k, (im, lab, spacing, origin) in enumerate(_dataset_cases()):
        d = tmp_path / ('case%d' % k)
        os.makedirs(d)
        write_image(Image3d(im * 100, spacing, origin), str(d / 'im.mha'), True)
        write_image(Image3d(lab.astype(np.int8), spacing, origin), str(d / 'seg.mha'), True)
        lines += [str(d / 'im.mha'), str(d / 'seg.mha')]
    with open(str(tmp_path / 'train.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    ds = SegmentationDataset(str(tmp_path / 'train.txt'), 3, [1.0, 1.0, 1.0], [32, 32, 16], 'HYBRID', [4, 4, 4], [0.9, 1.1], 'LINEAR',
                             [FixedNormalizer(0.0, 100.0, True)])
    for workers in (0, 2):
        loader = DataLoader(ds, sampler=EpochConcateSampler(ds, 2), batch_size=2, num_workers=workers, pin_memory=False)
        n = 0
        for crops, masks, frames, names in loader:
            assert tuple(crops.shape) == (2, 1, 16, 32, 32) and crops.dtype == torch.float32
            assert tuple(masks.shape) == (2, 1, 16, 32, 32) and masks.dtype == torch.float32
            assert tuple(frames.shape) == (2, 15) and len(names) == 2 and names[0].startswith('case')
            assert float(crops.abs().max()) <= 1.0 + 1e-6                 # normaliser clip
            assert set(np.unique(masks.numpy())) <= {0.0, 1.0, 2.0}
            n += 1
        assert n == 3


def test_samplers_draw_the_reference_index_streams():
    """dataloader/sampler.py:6-79: same index streams as the unmodified reference under python's `random`
    (tests/golden/samplers.json); the distributed sampler is DistributedSampler's shuffle per epoch, rank-sharded."""
    import json
    import random
    from torch.utils.data.distributed import DistributedSampler
    from segmentation3d.dataloader.sampler import EpochConcateDistributedSampler, EpochConcateSampler, EpochConcateSamplerResume
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'samplers.json')))
    random.seed(5)
    s = EpochConcateSampler(list(range(7)), 3)
    assert list(s) == gold['concat_n7_e3_seed5'] and len(s) == 21
    s = EpochConcateSamplerResume(list(range(5)), 3, 2)
    assert list(s) == gold['resume_n5_e3_from2'] and len(s) == 15
    data = list(range(11))
    streams = []
    for r in range(2):
        s = EpochConcateDistributedSampler(data, 3, 0, rank=r, world_size=2, seed=9)
        got = list(s)
        ref = []
        base = DistributedSampler(data, num_replicas=2, rank=r, seed=9)
        for e in range(3):
            base.set_epoch(e)
            ref += list(base)
        assert got == ref and len(s) == len(got) == 18          # ceil(11 / 2) per epoch: every rank runs the same number of steps
        streams.append(got)
    for e in range(3):
        both = streams[0][6 * e:6 * e + 6] + streams[1][6 * e:6 * e + 6]
        assert set(both) == set(data)                           # every sample once per epoch (plus one padding repeat)


def test_image_tools_helpers(tmp_path):
    """host helpers with the reference's names: get_bounding_box, copy_image, percentiles, frames, save_intermediate_results."""
    from segmentation3d.utils.image3d import read_image
    from segmentation3d.utils.image_tools import (copy_image, get_bounding_box, get_image_frame, percentiles,
                                                  save_intermediate_results, set_image_frame)
    m = np.zeros((10, 12, 14), np.int8)
    m[2:5, 3:9, 4:6] = 1
    m[7, 1, 12] = 2
    assert get_bounding_box(Image3d(m), None) == ([4, 1, 2], [13, 9, 8])
    assert get_bounding_box(Image3d(m), [1]) == ([4, 3, 2], [6, 9, 5])
    assert get_bounding_box(Image3d(m), [3]) == (None, None)
    from segmentation3d.core import seg_infer
    assert seg_infer.get_bounding_box is get_bounding_box                      # the engine's cascade uses the same function
    rng = np.random.default_rng(0)
    src = Image3d(rng.standard_normal((6, 6, 6)).astype(np.float32), (1.0, 1.0, 1.0), (4.0, 3.0, 2.0))
    tgt = Image3d(np.zeros((10, 12, 14), np.float32))
    bs, be = [5.0, 4.0, 3.0], [9, 8, 7]
    out = copy_image(src, bs, be, tgt).to_numpy()
    assert bs == [5, 4, 3] and isinstance(bs[0], int)
    assert np.array_equal(out[3:7, 4:8, 5:9], src.to_numpy()[1:5, 1:5, 1:5]) and out.sum() == out[3:7, 4:8, 5:9].sum()
    assert np.allclose(percentiles(Image3d(m), [50, 100]), np.percentile(m, [50, 100]))
    im = Image3d(np.zeros((2, 3, 4), np.float32), (0.5, 0.6, 0.7), (1.0, 2.0, 3.0))
    fr = get_image_frame(im)
    assert fr.dtype == np.float32 and np.allclose(fr[:6], [0.5, 0.6, 0.7, 1.0, 2.0, 3.0])
    im2 = Image3d(np.zeros((2, 3, 4), np.float32))
    set_image_frame(im2, fr)
    assert np.allclose(im2.GetSpacing(), im.GetSpacing()) and np.allclose(im2.GetOrigin(), im.GetOrigin())
    crops, masks = torch.randn(2, 1, 4, 5, 6), torch.randint(0, 3, (2, 1, 4, 5, 6)).float()
    outputs = torch.softmax(torch.randn(2, 3, 4, 5, 6), 1)
    frames = torch.from_numpy(np.stack([fr, fr]))
    save_intermediate_results([1], crops, masks, outputs, frames, ['caseA_im.mha', 'caseB_im.mha'], str(tmp_path / 'batch_0'))
    d = tmp_path / 'batch_0' / 'caseB_im.mha'
    assert sorted(os.listdir(d)) == ['batch_1_crop_0.nii.gz', 'batch_1_mask.nii.gz', 'batch_1_output_0.nii.gz',
                                     'batch_1_output_1.nii.gz', 'batch_1_output_2.nii.gz']
    back = read_image(str(d / 'batch_1_mask.nii.gz'))
    assert back.to_numpy().dtype == np.int32 and np.array_equal(back.to_numpy(), masks[1, 0].numpy().astype(np.int32))
    assert np.allclose(back.GetSpacing(), (0.5, 0.6, 0.7), atol=1e-6) and np.allclose(back.GetOrigin(), (1.0, 2.0, 3.0), atol=1e-5)


def _fake_engine(monkeypatch, log):
    """segmentation() with the GPU parts replaced by host stand-ins: the case loop, list readers, sharding, I/O overlap and
    file layout are the code under test"""
    from segmentation3d.core import seg_infer
    from segmentation3d.utils.attrdict import AttrDict

    def load_models(model_folder, gpu_id=0):
        cfg = AttrDict({'general': {'single_scale': 'fine'}, 'fine': {}, 'coarse': {}})
        m = AttrDict()
        dict.__setitem__(m, 'infer_cfg', cfg)
        dict.__setitem__(m, 'fine_model', {'out_channels': 2})
        dict.__setitem__(m, 'coarse_model', None)
        return m

    def segmentation_volume(model, cfg, image, bs, be, use_gpu):
        a = image.to_numpy()
        log.append(float(a.flat[0]))
        p1 = (1.0 / (1.0 + np.exp(-a))).astype(np.float32)
        probs = []
        for p in (1.0 - p1, p1):
            im = Image3d(p)
            im.CopyInformation(image)
            probs.append(im)
        mask = Image3d((p1 > 0.5).astype(np.int8))
        mask.CopyInformation(image)
        return probs, mask
    monkeypatch.setattr(seg_infer, 'load_models', load_models)
    monkeypatch.setattr(seg_infer, 'segmentation_volume', segmentation_volume)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    return seg_infer


def test_segmentation_case_loop_with_background_io(tmp_path, monkeypatch):
    """core/seg_infer.py:353-493: the case loop writes the same files, in the same layout, whether reads / writes overlap
    the computation (default) or run strictly in series (SEG3D_IO_THREADS=0); a missing case raises when it is reached,
    after the earlier cases have been written completely."""
    from segmentation3d.utils.image3d import read_image, write_image
    log = []
    seg_infer = _fake_engine(monkeypatch, log)
    rng = np.random.default_rng(1)
    lines = ['5']
    for k in range(5):
        a = rng.standard_normal((6, 8, 10)).astype(np.float32)
        a.flat[0] = k
        write_image(Image3d(a, (1.0, 1.0, 2.0), (k, 0.0, 0.0)), str(tmp_path / ('im%d.mha' % k)), True)
        lines.append('case%d %s' % (k, tmp_path / ('im%d.mha' % k)))
    with open(str(tmp_path / 'test.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    outs = {}
    for threads in ('2', '0'):
        monkeypatch.setenv('SEG3D_IO_THREADS', threads)
        del log[:]
        out = tmp_path / ('out' + threads)
        masks = seg_infer.segmentation(str(tmp_path / 'test.txt'), str(tmp_path), str(out), 'seg.mha', 0, True, True, True, True)
        assert log == [0.0, 1.0, 2.0, 3.0, 4.0] and len(masks) == 5                  # cases in list order
        assert sorted(os.listdir(out)) == ['case%d' % k for k in range(5)]
        for k in range(5):
            d = out / ('case%d' % k)
            assert sorted(os.listdir(d)) == ['mean_prob_0.mha', 'mean_prob_1.mha', 'org.mha', 'seg.mha']
            seg = read_image(str(d / 'seg.mha'))
            assert seg.to_numpy().dtype == np.int8 and np.array_equal(seg.to_numpy(), masks[k].to_numpy())
            assert np.allclose(seg.GetOrigin(), (k, 0.0, 0.0)) and np.allclose(seg.GetSpacing(), (1.0, 1.0, 2.0))
            outs[(threads, k)] = {n: open(str(d / n), 'rb').read() for n in os.listdir(d)}
    for k in range(5):
        assert outs[('2', k)] == outs[('0', k)]                                     # byte-identical files
    # a listed file that does not exist is refused when the list is read, like the reference (core/seg_infer.py:41-42)
    lines[3] = 'case2 %s' % (tmp_path / 'missing.mha')
    with open(str(tmp_path / 'missing.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    with pytest.raises(ValueError, match='image not exist'):
        seg_infer.segmentation(str(tmp_path / 'missing.txt'), str(tmp_path), str(tmp_path / 'none'), 'seg.mha', 0, False, True, False, False)
    # a case that cannot be decoded (truncated pixel data): raised at its turn, earlier cases complete on disk
    with open(str(tmp_path / 'truncated.mha'), 'wb') as f:
        f.write(b'ObjectType = Image\nNDims = 3\nDimSize = 4 4 4\nElementType = MET_FLOAT\nElementDataFile = LOCAL\n' + b'\0' * 10)
    lines[3] = 'case2 %s' % (tmp_path / 'truncated.mha')
    with open(str(tmp_path / 'bad.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    for threads in ('2', '0'):
        monkeypatch.setenv('SEG3D_IO_THREADS', threads)
        out = tmp_path / ('bad' + threads)
        with pytest.raises(ValueError, match='buffer is smaller'):
            seg_infer.segmentation(str(tmp_path / 'bad.txt'), str(tmp_path), str(out), 'seg.mha', 0, False, True, False, False)
        assert sorted(os.listdir(out)) == ['case0', 'case1']
        assert read_image(str(out / 'case1' / 'seg.mha')).GetSize() == (10, 8, 6)
    # torchrun-style case sharding goes through the same loop
    monkeypatch.setenv('SEG3D_IO_THREADS', '2')
    monkeypatch.setenv('WORLD_SIZE', '2'); monkeypatch.setenv('RANK', '1'); monkeypatch.setenv('LOCAL_RANK', '1')
    del log[:]
    seg_infer.segmentation(str(tmp_path / 'test.txt'), str(tmp_path), str(tmp_path / 'sh'), 'seg.mha', 0, False, True, False, False)
    assert log == [1.0, 3.0] and sorted(os.listdir(tmp_path / 'sh')) == ['case1', 'case3']


def test_dataset_volume_cache_changes_nothing_but_the_reads(tmp_path, monkeypatch):
    """SEG3D_DATASET_CACHE_MB: decoded volumes are kept per process; items are identical with and without the cache, and
    the byte bound evicts the least recently used volume."""
    from segmentation3d.dataloader.dataset import SegmentationDataset
    from segmentation3d.utils.image3d import write_image
    lines = ['3']
    for k, (im, lab, spacing, origin) in enumerate(_dataset_cases()):
        d = tmp_path / ('case%d' % k)
        os.makedirs(d)
        write_image(Image3d(im, spacing, origin), str(d / 'im.mha'), True)
        write_image(Image3d(lab, spacing, origin), str(d / 'seg.mha'), True)
        lines += [str(d / 'im.mha'), str(d / 'seg.mha')]
    with open(str(tmp_path / 'train.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    items = {}
    for mb in ('0', '2048', '0.8'):
        monkeypatch.setenv('SEG3D_DATASET_CACHE_MB', mb)
        ds = SegmentationDataset(str(tmp_path / 'train.txt'), 3, [1.0, 1.0, 1.0], [16, 16, 16], 'HYBRID', [3, 3, 3], [0.9, 1.1], 'LINEAR', [None])
        np.random.seed(3)
        items[mb] = [ds[i] for i in (0, 1, 0, 2, 1, 0)]
        if mb == '0':
            assert ds._cache.hits == 0 and ds._cache.misses == 12 and not ds._cache.items
        elif mb == '2048':
            assert ds._cache.misses == 6 and ds._cache.hits == 6
        else:                                   # 0.8 MB holds two cases (four 0.17 MB volumes): case 1, then case 0, get evicted
            assert 0 < ds._cache.bytes <= 0.8 * (1 << 20) and ds._cache.misses == 10 and ds._cache.hits == 2
    for a, b, c in zip(items['0'], items['2048'], items['0.8']):
        for other in (b, c):
            assert torch.equal(a[0], other[0]) and torch.equal(a[1], other[1]) and np.array_equal(a[2], other[2]) and a[3] == other[3]


def test_trilinear_restatements_agree_with_scipy_map_coordinates():
    """SimpleITK is absent, so the ITK restatements cannot be pinned against ITK itself; scipy.ndimage.map_coordinates
    (order=1, mode='nearest' = both neighbours clamped) is an independent implementation of the same interpolation: the
    oracle's resample_grid (inference side) and the product's crop_image (training side) must agree with it wherever the
    continuous index is inside the volume, and write the default value elsewhere."""
    from scipy import ndimage
    from segmentation3d.utils.image_tools import crop_image
    rng = np.random.default_rng(8)
    src = rng.standard_normal((18, 22, 26)).astype(np.float32)
    sp_in, sp_out, size = (0.9, 1.2, 2.0), (1.3, 0.7, 1.1), [20, 30, 28]
    cz, cy, cx = [np.arange(size[a], dtype=np.float64) * (sp_out[a] / sp_in[a]) for a in (2, 1, 0)]
    grid = np.meshgrid(cz, cy, cx, indexing='ij')
    ref = ndimage.map_coordinates(src.astype(np.float64), grid, order=1, mode='nearest')
    inside = (grid[0] < 18 - 0.5) & (grid[1] < 22 - 0.5) & (grid[2] < 26 - 0.5)
    got = orz.resample_grid(src, sp_in, size, sp_out, 'LINEAR', -7.0)
    assert np.abs(got[inside] - ref[inside]).max() <= 1e-5 and (got[~inside] == -7.0).all() and inside.mean() > 0.3
    nn = orz.resample_grid(src, sp_in, size, sp_out, 'NN', 0.0)
    ref_nn = src[np.floor(grid[0] + 0.5).astype(int).clip(0, 17), np.floor(grid[1] + 0.5).astype(int).clip(0, 21),
                 np.floor(grid[2] + 0.5).astype(int).clip(0, 25)]
    assert np.array_equal(nn[inside], ref_nn[inside])
    # crop with an offset origin (continuous indices below 0 and beyond the far faces)
    origin = (4.0, -3.0, 10.0)
    img = Image3d(src, sp_in, origin)
    csz, csp, center = [24, 20, 16], (0.8, 1.0, 1.7), (8.0, 20.0, 40.0)
    out = crop_image(img, center, csz, csp, 'LINEAR').to_numpy()
    o = [center[a] - csz[a] * csp[a] / 2.0 + csp[a] / 2.0 for a in range(3)]
    c = [(o[a] - origin[a]) / sp_in[a] + np.arange(csz[a], dtype=np.float64) * (csp[a] / sp_in[a]) for a in range(3)]
    grid = np.meshgrid(c[2], c[1], c[0], indexing='ij')
    ref = ndimage.map_coordinates(src.astype(np.float64), grid, order=1, mode='nearest')
    n_in = (18, 22, 26)
    inside = np.ones(ref.shape, bool)
    for a in range(3):
        inside &= (grid[a] >= -0.5) & (grid[a] < n_in[a] - 0.5)
    assert 0.2 < inside.mean() < 0.95
    assert np.abs(out[inside] - ref[inside]).max() <= 1e-5 and (out[~inside] == 0).all()


def test_patch_grid_equals_oracle_on_random_configurations():
    """image_partition_by_fixed_size (utils/image_tools.py:163-218) against the oracle's restatement (itself pinned to 14
    reference grids) on 400 random volumes / spacings / boxes / partition sizes and strides, in-place bbox update included."""
    from oracle import sliding_window as osw
    from segmentation3d.utils.image_tools import image_partition_by_fixed_size
    rng = np.random.default_rng(2024)
    checked = 0
    for _ in range(400):
        size = [int(16 * rng.integers(1, 12)) for _ in range(3)]
        spacing = [float(rng.choice([0.4, 0.5, 0.8, 1.0, 1.25, 2.0])) for _ in range(3)]
        if rng.random() < 0.5:
            bs, be = [0, 0, 0], list(size)
        else:
            bs = [int(rng.integers(0, size[a] - 1)) for a in range(3)]
            be = [int(rng.integers(bs[a] + 1, size[a] + 1)) for a in range(3)]
        psize = [float(rng.choice([16, 24, 32, 48, 64, 96, 51.2, 89.6])) for _ in range(3)]
        pstride = [float(max(4.0, psize[a] * rng.choice([0.25, 0.5, 0.75, 1.0, 1.5]))) for a in range(3)]
        im = Image3d(np.zeros((1, 1, 1), np.float32), spacing)
        im.GetSize = lambda s=tuple(size): s
        b1, e1, b2, e2 = list(bs), list(be), list(bs), list(be)
        try:
            ref = osw.partition_grid(size, spacing, b2, e2, list(psize), list(pstride), 16)
        except AssertionError:
            with pytest.raises(AssertionError):
                image_partition_by_fixed_size(im, b1, e1, list(psize), list(pstride), 16)
            continue
        got = image_partition_by_fixed_size(im, b1, e1, list(psize), list(pstride), 16)
        assert [list(map(int, s)) for s in got[0]] == [list(map(int, s)) for s in ref[0]]
        assert [list(map(int, s)) for s in got[1]] == [list(map(int, s)) for s in ref[1]]
        assert b1 == b2 and e1 == e2
        checked += 1
    assert checked > 300


def test_async_writer_and_prefetch_surface_errors(tmp_path):
    """utils/image3d.py: a failed background write is raised by close() after every other file has been written; a failed
    read is raised when its own case is reached; both helpers degrade to serial calls with zero threads."""
    from segmentation3d.utils.image3d import AsyncImageWriter, prefetch_images, read_image, write_image
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    for threads in (0, 3):
        w = AsyncImageWriter(threads)
        w.write(Image3d(a), str(tmp_path / ('ok%d_0.mha' % threads)), True)
        if threads:
            w.write(Image3d(a), str(tmp_path / 'x.tiff'), True)                  # unsupported extension -> ValueError in the worker
        else:
            with pytest.raises(ValueError):
                w.write(Image3d(a), str(tmp_path / 'x.tiff'), True)
        w.write(Image3d(a + 1), str(tmp_path / ('ok%d_1.mha' % threads)), False)
        if threads:
            with pytest.raises(ValueError):
                w.close()
        else:
            w.close()
        assert np.array_equal(read_image(str(tmp_path / ('ok%d_0.mha' % threads))).to_numpy(), a)
        assert np.array_equal(read_image(str(tmp_path / ('ok%d_1.mha' % threads))).to_numpy(), a + 1)
        w.close()                                                                 # idempotent
    paths = []
    for k in range(4):
        write_image(Image3d(a + k), str(tmp_path / ('p%d.mha' % k)))
        paths.append(str(tmp_path / ('p%d.mha' % k)))
    paths.insert(2, str(tmp_path / 'nope.mha'))
    for enabled, depth in ((True, 3), (True, 1), (False, 3)):
        seen = []
        with pytest.raises(FileNotFoundError):
            for img, waited in prefetch_images(paths, np.float32, enabled=enabled, depth=depth):
                seen.append(float(img.to_numpy().flat[0]))
                assert waited >= 0.0
        assert seen == [0.0, 1.0]
    assert [float(i.to_numpy().flat[0]) for i, _ in prefetch_images(paths[:2] + paths[3:], np.float32, depth=2)] == [0.0, 1.0, 2.0, 3.0]


def test_segmentation_cascade_orchestration(tmp_path, monkeypatch, capsys):
    """core/seg_infer.py:428-444 (single_scale = 'DISABLE'): the coarse model sees the whole image, the bounding box of its
    mask (all non-zero labels, end exclusive) is handed to the fine model, whose result is what gets written."""
    from segmentation3d.core import seg_infer
    from segmentation3d.utils.attrdict import AttrDict
    from segmentation3d.utils.image3d import read_image, write_image
    seen = []

    def load_models(model_folder, gpu_id=0):
        m = AttrDict()
        dict.__setitem__(m, 'infer_cfg', AttrDict({'general': {'single_scale': 'DISABLE'}, 'fine': {'tag': 'fine'}, 'coarse': {'tag': 'coarse'}}))
        dict.__setitem__(m, 'fine_model', {'out_channels': 2, 'tag': 'fine'})
        dict.__setitem__(m, 'coarse_model', {'out_channels': 2, 'tag': 'coarse'})
        return m

    def segmentation_volume(model, cfg, image, bs, be, use_gpu):
        seen.append((model['tag'], cfg['tag'], bs, be))
        a = image.to_numpy()
        lab = np.zeros(a.shape, np.int8)
        if model['tag'] == 'coarse':
            lab[3:7, 2:9, 5:11] = 1
            lab[4, 4, 12] = 2
        else:
            lab[bs[2]:be[2], bs[1]:be[1], bs[0]:be[0]] = 1
        probs = [Image3d((lab == c).astype(np.float32)) for c in range(2)]
        mask = Image3d(lab)
        mask.CopyInformation(image)
        return probs, mask
    monkeypatch.setattr(seg_infer, 'load_models', load_models)
    monkeypatch.setattr(seg_infer, 'segmentation_volume', segmentation_volume)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    write_image(Image3d(np.zeros((10, 12, 16), np.float32)), str(tmp_path / 'im.mha'))
    masks = seg_infer.segmentation(str(tmp_path / 'im.mha'), str(tmp_path), str(tmp_path / 'out'), 'seg.mha', 1, True, True, False, False)
    assert seen == [('coarse', 'coarse', None, None), ('fine', 'fine', [5, 2, 3], [13, 9, 7])]
    out = read_image(str(tmp_path / 'out' / 'im.mha' / 'seg.mha')).to_numpy()
    assert out.sum() == 8 * 7 * 4 and np.array_equal(out, masks[0].to_numpy())
    assert 'Fine segmentation (bbox ratio: %.2f%%)' % (100 * (8 / 16) * (7 / 12) * (4 / 10)) in capsys.readouterr().out
