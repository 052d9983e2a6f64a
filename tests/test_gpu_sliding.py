"""GPU parity of the device sliding-window engine against the reference goldens and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import init as oinit
from oracle import sliding_window as osw
from oracle.metrics import parity_report

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


def seeded_input(seed, shape, kind='noise'):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    if kind == 'smooth':
        lo = torch.randn((shape[0], shape[1]) + tuple(max(2, s // 8) for s in shape[2:]), generator=g)
        x = torch.nn.functional.interpolate(lo, size=shape[2:], mode='trilinear', align_corners=False) * 2 + 0.3 * x
    return x


def build_model(arch, cout, sd, mode, norm):
    import importlib
    from segmentation3d.core.seg_infer import make_model
    mod = importlib.import_module('segmentation3d.network.' + arch)
    net = mod.SegmentationNet(1, cout)
    net.load_state_dict(sd)
    net.b200_mode = mode
    net = net.cuda().eval()
    return make_model(net, [1.0, 1.0, 1.0], norm, batch=3)


def _run_golden_case(z, case, mode):
    from segmentation3d.core.seg_infer import segmentation_volume
    from segmentation3d.utils.image3d import Image3d
    name, arch, cout, wseed, aseed, size, psize, pstride, norm, vseed, scale = case
    sd = oinit.init_state_dict(arch, 1, cout, wseed)
    if aseed is not None:
        sd = oinit.randomize_affine(sd, aseed)
    vol = (seeded_input(vseed, (1, 1, size[2], size[1], size[0]), 'smooth')[0, 0].numpy() * scale).astype(np.float32)
    nd = {'type': 0, 'mean': norm[1], 'stddev': norm[2], 'clip': norm[3]} if norm[0] == 'fixed' else {'type': 1, 'clip_sigma': norm[1]}
    model = build_model(arch, cout, sd, mode, nd)
    cfg = {'partition_type': 'SIZE', 'partition_size': psize, 'partition_stride': pstride,
           'pick_largest_cc': False, 'remove_small_cc': 0}
    probs_im, mask_im = segmentation_volume(model, cfg, Image3d(vol), None, None, True)
    probs = np.stack([p.to_numpy() for p in probs_im], 0)
    mask = mask_im.to_numpy()
    assert mask.dtype == np.int8
    # mask must be the first-argmax of the returned probabilities
    assert np.array_equal(mask, osw.argmax_first(probs))
    return probs, mask


@pytest.mark.parametrize('mode', ['fp32', 'fp32x'])
def test_segmentation_volume_matches_reference_golden(mode):
    """strict bars (CUDA-core fp32 and split-operand tensor-core mode) on the reference's own outputs: 32^3 patches, fixed and
    adaptive normaliser, overlapping strides, VNet and VBNet C=5"""
    z = np.load(os.path.join(G, 'sliding_window.npz'))
    for case in json.loads(str(z['meta'])):
        name = case[0]
        probs, mask = _run_golden_case(z, case, mode)
        rep = parity_report(z[name + '_probs'], probs)
        agree_mask = float((mask == z[name + '_mask']).mean())
        print(name, mode, rep, 'mask agreement vs reference mask %.5f' % agree_mask)
        assert rep['max_abs'] <= 1e-3, (name, rep)
        assert agree_mask >= 0.999, name
        assert min(rep['dice']) >= 0.999, (name, rep)


@pytest.mark.parametrize('mode', ['fp32x', 'fp16'])
def test_segmentation_volume_64_patches_meets_bars(mode):
    """BASELINE.json bars on reference outputs with 64^3 patches (tests/golden/sliding_window_64.npz: overlapping stride 48
    with a fixed normaliser, tiled stride 64 with the adaptive normaliser): fp32x <= 1e-3; fp16 <= 1e-2, label agreement
    >= 99.9 %, per-class Dice >= 0.999 against the REFERENCE's mask.  Probabilities are committed on every 2nd voxel."""
    from oracle.metrics import cal_dsc
    z = np.load(os.path.join(G, 'sliding_window_64.npz'))
    for case in json.loads(str(z['meta'])):
        name, cout = case[0], case[2]
        probs, mask = _run_golden_case(z, case, mode)
        max_abs = float(np.abs(probs[:, ::2, ::2, ::2] - z[name + '_probs']).max())
        agree = float((mask == z[name + '_mask']).mean())
        dice = [float(cal_dsc(z[name + '_mask'], mask, c, 1)[0]) for c in range(cout)]
        print(name, mode, 'max|dp| %.3g, label agreement %.5f, per-class Dice %s' % (max_abs, agree, dice))
        assert max_abs <= (1e-3 if mode == 'fp32x' else 1e-2), (name, max_abs)
        assert agree >= 0.999 and min(dice) >= 0.999, (name, agree, dice)


def test_patch_grid_count_and_blend_kernels():
    """blend + finalize kernels against the oracle's numpy accumulate on random probabilities."""
    from segmentation3d._b200 import lib as L
    from segmentation3d._b200.sliding import axis_counts
    L.load()
    size = [64, 48, 80]
    starts, ends = osw.partition_grid(size, [1, 1, 1], [0, 0, 0], list(size), [32, 32, 32], [16, 16, 16], 16)
    C, n = 3, len(starts)
    g = torch.Generator().manual_seed(1)
    probs = torch.rand((n, C, 32, 32, 32), generator=g)
    acc_ref = np.zeros((C, size[2], size[1], size[0]), np.float32)
    for i, (s, e) in enumerate(zip(starts, ends)):
        acc_ref[:, s[2]:e[2], s[1]:e[1], s[0]:e[0]] += probs[i].numpy()
    cnt = osw.overlap_count_axes(size, starts, ends)
    cx, cy, cz = axis_counts(size, starts, ends)
    assert np.array_equal(cnt, (cz[:, None, None] * cy[None, :, None] * cx[None, None, :]).astype(np.float32))
    ref = acc_ref * (np.float32(1.0) / cnt)
    acc = torch.zeros((C, size[2], size[1], size[0]), device='cuda')
    sd = torch.tensor(np.asarray(starts, np.int32), device='cuda')
    pd = probs.cuda()
    # both code paths: four voxels per vector atomic (every start x is a multiple of 4 here) and the scalar one
    acc1 = torch.zeros_like(acc)
    L.call('seg3d_blend_accumulate', L.ptr(pd), n, C, 32, 32, 32, L.ptr(sd), L.ptr(acc1), size[2], size[1], size[0], 0, L.stream_ptr())
    assert all(s[0] % 4 == 0 for s in starts)
    L.call('seg3d_blend_accumulate', L.ptr(pd), n, C, 32, 32, 32, L.ptr(sd), L.ptr(acc), size[2], size[1], size[0], 1, L.stream_ptr())
    torch.cuda.synchronize()
    assert float((acc - acc1).abs().max()) <= 1e-5
    mask = torch.empty((size[2], size[1], size[0]), dtype=torch.int8, device='cuda')
    cxd, cyd, czd = [torch.tensor(c, device='cuda') for c in (cx, cy, cz)]
    L.call('seg3d_blend_finalize_argmax', L.ptr(acc), C, size[2], size[1], size[0], L.ptr(cxd), L.ptr(cyd), L.ptr(czd), L.ptr(mask), L.stream_ptr())
    torch.cuda.synchronize()
    got = acc.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-5          # float add order differs across overlapping patches
    assert np.array_equal(mask.cpu().numpy(), osw.argmax_first(got))


def test_patch_gather_normalisers_bit_exact_fixed():
    from segmentation3d._b200 import lib as L
    L.load()
    g = torch.Generator().manual_seed(2)
    vol = (torch.randn((40, 48, 56), generator=g) * 300).float()
    starts = [[3, 5, 7], [24, 16, 8]]
    sd = torch.tensor(np.asarray(starts, np.int32), device='cuda')
    vd = vol.cuda()
    out = torch.empty((2, 32, 32, 32), device='cuda')
    L.call('seg3d_patch_gather', L.ptr(vd), 40, 48, 56, L.ptr(sd), 2, 32, 32, 32, L.NORM_FIXED, 50.0, 200.0, 1, -1.0, 1.0,
           None, L.F32, L.ptr(out), L.stream_ptr())
    torch.cuda.synchronize()
    for i, s in enumerate(starts):
        ref = osw.normalize_fixed(vol.numpy()[s[2]:s[2] + 32, s[1]:s[1] + 32, s[0]:s[0] + 32], 50.0, 200.0, True)
        assert np.array_equal(out[i].cpu().numpy(), ref)         # float32 sub + div: bit-exact
    # adaptive: statistics are reduced in a different order -> tolerance
    stats = torch.zeros((2, 2), dtype=torch.float64, device='cuda')
    L.call('seg3d_patch_stats', L.ptr(vd), 40, 48, 56, L.ptr(sd), 2, 32, 32, 32, L.ptr(stats), L.stream_ptr())
    L.call('seg3d_patch_gather', L.ptr(vd), 40, 48, 56, L.ptr(sd), 2, 32, 32, 32, L.NORM_ADAPTIVE, 0.0, 1.0, 1, -2.5, 2.5,
           L.ptr(stats), L.F32, L.ptr(out), L.stream_ptr())
    torch.cuda.synchronize()
    for i, s in enumerate(starts):
        ref = osw.normalize_adaptive(np.ascontiguousarray(vol.numpy()[s[2]:s[2] + 32, s[1]:s[1] + 32, s[0]:s[0] + 32]), 2.5)
        assert np.abs(out[i].cpu().numpy() - ref).max() <= 1e-5


@pytest.mark.parametrize('dt_name', ['F16', 'BF16', 'F32'])
def test_patch_gather_rows_writes_the_padded_layout(dt_name):
    """seg3d_patch_gather_rows (the row-padded input of the Toeplitz input block): every voxel lands at column x_off + x with the
    value seg3d_patch_gather stores, and the padding columns are zero (2-byte types: written as zeros by the word-wise kernel;
    fp32: left as the caller initialised them)."""
    from segmentation3d._b200 import lib as L
    L.load()
    dt = getattr(L, dt_name)
    tdt = L.TORCH_DTYPE[dt]
    g = torch.Generator().manual_seed(4)
    vol = (torch.randn((40, 48, 56), generator=g) * 300).float()
    starts = [[3, 5, 7], [24, 16, 8], [0, 0, 0]]
    sd = torch.tensor(np.asarray(starts, np.int32), device='cuda')
    vd = vol.cuda()
    pz, py, px = 24, 32, 32
    pitch, off = px + L.CIN1_PAD, L.CIN1_LEFT
    dense = torch.empty((3, pz, py, px), dtype=tdt, device='cuda')
    L.call('seg3d_patch_gather', L.ptr(vd), 40, 48, 56, L.ptr(sd), 3, pz, py, px, L.NORM_FIXED, 50.0, 200.0, 1, -1.0, 1.0,
           None, dt, L.ptr(dense), L.stream_ptr())
    fill = 0.0 if dt == L.F32 else 7.0            # the word-wise kernel must overwrite the padding of the 2-byte layouts with zeros
    rows = torch.full((3, pz, py, pitch), fill, dtype=tdt, device='cuda')
    L.call('seg3d_patch_gather_rows', L.ptr(vd), 40, 48, 56, L.ptr(sd), 3, pz, py, px, L.NORM_FIXED, 50.0, 200.0, 1, -1.0, 1.0,
           None, dt, L.ptr(rows), pitch, off, L.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(rows[..., off:off + px], dense)
    assert float(rows[..., :off].float().abs().max()) == 0 and float(rows[..., off + px:].float().abs().max()) == 0


@pytest.mark.parametrize('size,psize,pstride', [((64, 48, 80), 32, 16), ((48, 48, 96), 32, 32)], ids=['overlap', 'tiled'])
def test_host_path_progressive_finalize_equals_device_path(size, psize, pstride):
    """segmentation_volume_host (slab-wise upload, z-ordered patches, slab-wise finalize + mask copy-out) must give the
    same probabilities and the same mask as the device-resident pass: the blend is order independent only up to fp32
    atomics, so compare with a tight tolerance, and the mask against the first-argmax of its own probabilities."""
    from segmentation3d.core.seg_infer import segmentation_volume_device, segmentation_volume_host
    sd = oinit.init_state_dict('vnet', 1, 2, 3)
    model = build_model('vnet', 2, sd, 'fp32', {'type': 0, 'mean': 0.0, 'stddev': 1.0, 'clip': False})
    cfg = {'partition_type': 'SIZE', 'partition_size': [psize] * 3, 'partition_stride': [pstride] * 3}
    vol = seeded_input(5, (1, 1, size[2], size[1], size[0]), 'smooth')[0, 0].contiguous()
    acc_d, mask_d = segmentation_volume_device(model, cfg, vol.cuda(), batch=3)
    host_vol = torch.empty(vol.shape, dtype=torch.float32, pin_memory=True)
    host_vol.copy_(vol)
    host_mask = torch.full(vol.shape, 77, dtype=torch.int8).pin_memory()
    acc_h, hm = segmentation_volume_host(model, cfg, host_vol, host_mask, batch=3)
    torch.cuda.synchronize()
    assert hm is host_mask and not (host_mask == 77).any()
    assert (acc_h - acc_d).abs().max() <= 1e-5
    assert np.array_equal(host_mask.numpy(), osw.argmax_first(acc_h.cpu().numpy()))
    assert float((host_mask.cuda() == mask_d).float().mean()) >= 0.9999


@pytest.mark.parametrize('interp', ['LINEAR', 'NN'])
@pytest.mark.parametrize('case', [((40, 36, 20), (0.7, 0.8, 2.5), (1.0, 1.0, 1.0)), ((32, 32, 48), (1.0, 1.0, 1.0), (0.4, 0.6, 0.5)),
                                  ((33, 17, 9), (1.3, 0.9, 1.1), (1.0, 1.0, 1.0))], ids=['aniso-to-iso', 'upsample', 'odd'])
def test_resample_kernel_matches_oracle(case, interp):
    """seg3d_resample vs the ITK-semantics restatement in oracle/resample.py (double-precision coordinates and weights on
    both sides: the float32 results must agree to the last bit, up to the fused-multiply-add contraction of the GPU)."""
    from oracle import resample as orz
    from segmentation3d.utils.image3d import Image3d
    from segmentation3d.utils.image_tools import resample, resample_spacing
    size, sp_in, sp_out = case
    vol = seeded_input(21, (1, 1, size[2], size[1], size[0]), 'smooth')[0, 0].numpy().astype(np.float32) * 100.0
    ref, osz = orz.resample_spacing(vol, sp_in, sp_out, 16, interp)
    im = resample_spacing(Image3d(vol, sp_in, (3.0, -2.0, 10.0)), sp_out, 16, interp)
    got = im.to_numpy()
    assert list(im.GetSize()) == osz and im.GetSpacing() == tuple(sp_out) and im.GetOrigin() == (3.0, -2.0, 10.0)
    assert got.dtype == np.float32 and got.shape == ref.shape
    tol = 0.0 if interp == 'NN' else 2e-5 * float(np.abs(vol).max())
    assert np.abs(got - ref).max() <= tol
    # and back onto the original grid with a padding value (core/seg_infer.py:330-333)
    back_ref = orz.resample_grid(ref, sp_out, size, sp_in, 'LINEAR', 1.0)
    back = resample(Image3d(ref, sp_out), Image3d(vol, sp_in), 'LINEAR', 1.0).to_numpy()
    assert np.abs(back - back_ref).max() <= 2e-5 * float(np.abs(vol).max())


def test_segmentation_volume_resamples_anisotropic_scan():
    """Scan spacing != model spacing: resample -> segment -> resample back -> argmax, against the oracle pipeline."""
    from segmentation3d.core.seg_infer import segmentation_volume
    from segmentation3d.utils.image3d import Image3d
    sd = oinit.init_state_dict('vnet', 1, 2, 4)
    norm = {'type': 0, 'mean': 0.0, 'stddev': 1.0, 'clip': False}
    model = build_model('vnet', 2, sd, 'fp32', norm)
    model['spacing'] = [1.0, 1.0, 1.0]
    size, sp = (56, 44, 20), (0.8, 0.9, 2.0)                       # -> iso 48 x 48 x 48 after rounding up to multiples of 16
    vol = seeded_input(8, (1, 1, size[2], size[1], size[0]), 'smooth')[0, 0].numpy().astype(np.float32)
    cfg = {'partition_type': 'SIZE', 'partition_size': [32, 32, 32], 'partition_stride': [16, 16, 16],
           'pick_largest_cc': False, 'remove_small_cc': 0}
    probs_im, mask_im = segmentation_volume(model, cfg, Image3d(vol, sp), None, None, True)
    probs = np.stack([p.to_numpy() for p in probs_im], 0)
    mask = mask_im.to_numpy()
    ref_probs, ref_mask = osw.segmentation_volume_resampled(sd, vol, sp, [1.0, 1.0, 1.0], norm, 'LINEAR', 16,
                                                            partition_size=[32, 32, 32], partition_stride=[16, 16, 16],
                                                            double_forward=False, faithful_copies=False)
    assert probs.shape == ref_probs.shape == (2,) + vol.shape and mask.shape == vol.shape and mask.dtype == np.int8
    assert mask_im.GetSpacing() == sp and probs_im[0].GetSpacing() == sp
    assert np.abs(probs - ref_probs).max() <= 1e-3
    assert float((mask == ref_mask).mean()) >= 0.999
    assert np.array_equal(mask, osw.argmax_first(probs))


def _blob_mask(seed, shape, nlabels):
    """random multi-label mask with many components of assorted sizes (thresholded smooth noise)"""
    g = torch.Generator().manual_seed(seed)
    m = np.zeros(shape, dtype=np.int8)
    for lab in range(1, nlabels + 1):
        lo = torch.randn((1, 1) + tuple(max(2, s // 6) for s in shape), generator=g)
        f = torch.nn.functional.interpolate(lo, size=shape, mode='trilinear', align_corners=False)[0, 0].numpy()
        f = f + 0.35 * torch.randn(shape, generator=g).numpy()
        m[(f > 0.9) & (m == 0)] = lab
    return m


@pytest.mark.parametrize('shape,nlabels', [((40, 48, 56), 1), ((64, 33, 50), 3), ((16, 16, 16), 2)], ids=['one', 'three', 'small'])
def test_connected_component_filters_match_scipy(shape, nlabels):
    """seg3d_cc_filter (union-find, 26-connectivity) vs the scipy restatement of pick_largest_connected_component /
    remove_small_connected_component: exact equality of the filtered masks."""
    from oracle import postprocess as opp
    from segmentation3d.utils.image_tools import pick_largest_connected_component, remove_small_connected_component
    m = _blob_mask(3, shape, nlabels)
    labels = list(range(1, nlabels + 1))
    assert all((m == lab).sum() > 0 for lab in labels)
    got = pick_largest_connected_component(m, labels).to_numpy()
    assert got.dtype == np.int8 and np.array_equal(got, opp.pick_largest_connected_component(m, labels))
    for thr in (1, 5, 40, 10 ** 6):
        got = remove_small_connected_component(m, labels, thr).to_numpy()
        assert np.array_equal(got, opp.remove_small_connected_component(m, labels, thr)), thr
    # degenerate inputs: empty label, one voxel, a full volume (one component)
    z = np.zeros((8, 8, 8), dtype=np.int8)
    assert not pick_largest_connected_component(z, [1]).to_numpy().any()
    z[3, 4, 5] = 1
    assert np.array_equal(pick_largest_connected_component(z, [1]).to_numpy(), z)
    full = np.ones((8, 16, 24), dtype=np.int8)
    assert np.array_equal(pick_largest_connected_component(full, [1]).to_numpy(), full)


def test_connected_components_large_volume_timing():
    """512 x 512 x 400 mask (the benchmark volume): one big component + specks; result checked by invariants (the kept
    voxels are a subset of the label, the kept component is at least as large as any other run-length estimate)."""
    from segmentation3d.core.seg_infer import _cc_filter_device
    g = torch.Generator(device='cuda').manual_seed(0)
    Z, Y, X = 400, 512, 512
    lo = torch.randn((1, 1, 13, 16, 16), generator=g, device='cuda')
    f = torch.nn.functional.interpolate(lo, size=(Z, Y, X), mode='trilinear', align_corners=False)[0, 0]
    m = ((f > 0.5) | (torch.rand((Z, Y, X), generator=g, device='cuda') > 0.9995)).to(torch.int8)
    del f
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = _cc_filter_device(m, [1], 0)
    e0.record()
    out = _cc_filter_device(m, [1], 0)
    e1.record()
    torch.cuda.synchronize()
    kept, total = int(out.sum()), int(m.sum())
    print('cc 512x512x400: %.2f ms, kept %d of %d label voxels' % (e0.elapsed_time(e1), kept, total))
    assert 0 < kept <= total and bool(((out == 1) <= (m == 1)).all())
    small = _cc_filter_device(m, [1], kept)            # threshold = size of the largest: only it survives
    assert torch.equal(small, out)


@pytest.mark.parametrize('mode', ['fp32', 'fp16'])
def test_segmentation_volume_with_bounding_box_matches_reference_golden(mode):
    """The coarse->fine cascade's call (core/seg_infer.py:292-307,428-444) against the reference's own output: inside
    the visited box probabilities and labels match; outside it the reference holds NaN (0 * 1/0) with label 0, the
    product holds probability 0 with label 0."""
    from segmentation3d.core.seg_infer import segmentation_volume
    from segmentation3d.utils.image3d import Image3d
    z = np.load(os.path.join(G, 'cascade.npz'))
    m = json.loads(str(z['meta']))
    sd = oinit.randomize_affine(oinit.init_state_dict(m['arch'], 1, m['cout'], m['wseed']), m['aseed'])
    size = m['size']
    vol = (seeded_input(m['vseed'], (1, 1, size[2], size[1], size[0]), 'smooth')[0, 0].numpy() * m['scale']).astype(np.float32)
    nd = {'type': 0, 'mean': m['norm'][1], 'stddev': m['norm'][2], 'clip': m['norm'][3]}
    model = build_model(m['arch'], m['cout'], sd, mode, nd)
    cfg = {'partition_type': 'SIZE', 'partition_size': m['psize'], 'partition_stride': m['pstride'],
           'pick_largest_cc': False, 'remove_small_cc': 0}
    probs_im, mask_im = segmentation_volume(model, cfg, Image3d(vol), list(m['bbox_start']), list(m['bbox_end']), True)
    probs = np.stack([p.to_numpy() for p in probs_im], 0)
    mask = mask_im.to_numpy()
    visited = np.isfinite(z['probs'][0])
    err = float(np.abs(probs[:, visited] - z['probs'][:, visited]).max())
    agree = float((mask[visited] == z['mask'][visited]).mean())
    print('bbox', mode, 'max|dp| inside %.3g, label agreement inside %.5f' % (err, agree))
    assert err <= (1e-3 if mode == 'fp32' else 2e-2)
    assert agree >= (0.999 if mode == 'fp32' else 0.99)
    assert not probs[:, ~visited].any() and not mask[~visited].any()
