"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py        (build container only)

/root/reference is put on sys.path read-only (after tests/golden/ref_shims.install()).  Nothing
here is imported by the product or by the GPU-side tests; the fixtures it writes are.
Fixtures (all small):
  schema.json            state-dict keys/shapes of vnet(1,2), vbnet(1,5)       (network/vnet.py, vbnet.py)
  weights_sha256.json    sha256 of seeded kaiming/gaussian-initialised weights  (module/weight_init.py)
  forward.npz            reference net outputs on seeded inputs                 (SegmentationNet.forward)
  grids.json             image_partition_by_fixed_size outputs                  (utils/image_tools.py:163-218)
  loss.npz               Dice / focal values and gradients                      (loss/*.py)
  sliding_window.npz     core.seg_infer.segmentation_volume end to end          (core/seg_infer.py:249-350)
  sliding_window_64.npz  the same with 64^3 patches (probabilities on every 2nd voxel per axis, full masks)
  train_step.npz         one Adam step of the reference training loop body      (core/seg_train.py:119-127)
  loss_ce.npz            cross-entropy values and gradients on loss.npz's inputs (loss/cross_entropy_loss.py)
  dataset_sampling.json  crops the reference's SegmentationDataset requests under a seeded RNG  (dataloader/dataset.py:140-209)
  samplers.json          index streams of EpochConcateSampler / EpochConcateSamplerResume  (dataloader/sampler.py:6-53)
  cascade.npz            segmentation_volume restricted by a bounding box       (core/seg_infer.py:292-307,428-444)
"""
import copy
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

sitk = ref_shims.install()
sys.path.insert(0, '/root/reference')
sys.dont_write_bytecode = True

from segmentation3d.utils import image_tools as ref_it            # noqa: E402
from segmentation3d.core import seg_infer as ref_infer            # noqa: E402
from segmentation3d.loss.multi_dice_loss import MultiDiceLoss    # noqa: E402
from segmentation3d.loss.focal_loss import FocalLoss              # noqa: E402
from segmentation3d.loss.binary_dice_loss import BinaryDiceLoss  # noqa: E402
from segmentation3d.utils.normalizer import FixedNormalizer, AdaptiveNormalizer  # noqa: E402

torch.set_num_threads(8)


def make_net(arch, cin, cout, seed, mode='kaiming'):
    mod = importlib.import_module('segmentation3d.network.' + arch)
    torch.manual_seed(seed)
    net = mod.SegmentationNet(cin, cout)
    (mod.parameters_kaiming_init if mode == 'kaiming' else mod.parameters_gaussian_init)(net)
    return net.eval()


def sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def seeded_input(seed, shape, kind='noise'):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    if kind == 'smooth':   # CT-like: low-frequency field + noise
        lo = torch.randn((shape[0], shape[1]) + tuple(max(2, s // 8) for s in shape[2:]), generator=g)
        x = torch.nn.functional.interpolate(lo, size=shape[2:], mode='trilinear', align_corners=False) * 2 + 0.3 * x
    return x


def randomize_affine(sd, seed):
    """must mirror oracle/init.py:randomize_affine"""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if v.dim() == 1:
            if '.gn' in k or '_gn' in k:
                out[k] = (1.0 + 0.2 * torch.randn(v.shape, generator=g)) if k.endswith('.weight') \
                    else 0.1 * torch.randn(v.shape, generator=g)
            else:
                out[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            out[k] = v.clone()
    return out


def gen_schema_and_hashes():
    schema, hashes = {}, {}
    for arch, cin, cout in (('vnet', 1, 2), ('vbnet', 1, 5)):
        net = make_net(arch, cin, cout, 0)
        sd = net.state_dict()
        schema['%s_%d_%d' % (arch, cin, cout)] = [[k, list(v.shape)] for k, v in sd.items()]
        for seed in (0, 1):
            hashes['%s_%d_%d_seed%d_kaiming' % (arch, cin, cout, seed)] = sd_hash(make_net(arch, cin, cout, seed).state_dict())
        hashes['%s_%d_%d_seed0_gaussian' % (arch, cin, cout)] = sd_hash(make_net(arch, cin, cout, 0, 'gaussian').state_dict())
    json.dump(schema, open(os.path.join(HERE, 'schema.json'), 'w'), indent=0)
    json.dump(hashes, open(os.path.join(HERE, 'weights_sha256.json'), 'w'), indent=1)


FORWARD_CASES = [
    # name, arch, cin, cout, weight seed, affine seed (None = as initialised), input seed, shape [B,C,D,H,W], kind
    ('vnet_c2_32', 'vnet', 1, 2, 0, None, 100, (1, 1, 32, 32, 32), 'noise'),
    ('vnet_c2_aff_16x32x48', 'vnet', 1, 2, 1, 7, 101, (2, 1, 16, 32, 48), 'smooth'),
    ('vbnet_c5_32', 'vbnet', 1, 5, 0, None, 102, (1, 1, 32, 32, 32), 'noise'),
    ('vbnet_c5_aff_32x16x48', 'vbnet', 1, 5, 1, 8, 103, (1, 1, 32, 16, 48), 'smooth'),
]


def gen_forward():
    out = {}
    meta = []
    for name, arch, cin, cout, wseed, aseed, iseed, shape, kind in FORWARD_CASES:
        net = make_net(arch, cin, cout, wseed)
        if aseed is not None:
            net.load_state_dict(randomize_affine(net.state_dict(), aseed))
        x = seeded_input(iseed, shape, kind)
        with torch.no_grad():
            y = net(x)
        out[name] = y.numpy().astype(np.float32)
        meta.append([name, arch, cin, cout, wseed, aseed, iseed, list(shape), kind])
        print(name, tuple(y.shape), float(y.min()), float(y.max()))
    out['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, 'forward.npz'), **out)


GRID_CASES = [
    # size xyz, spacing, bbox_start, bbox_end (None = whole), partition_size mm, stride mm
    ([512, 512, 400], [1, 1, 1], None, None, [96, 96, 96], [96, 96, 96]),
    ([512, 512, 400], [1, 1, 1], None, None, [96, 96, 96], [48, 48, 48]),
    ([256, 256, 256], [1, 1, 1], None, None, [96, 96, 96], [96, 96, 96]),
    ([256, 256, 256], [1, 1, 1], None, None, [96, 96, 96], [48, 48, 48]),
    ([96, 96, 96], [1, 1, 1], None, None, [96, 96, 96], [96, 96, 96]),
    ([192, 192, 192], [1, 1, 1], None, None, [96, 96, 96], [96, 96, 96]),
    ([512, 512, 400], [0.4, 0.4, 0.4], None, None, [89.6, 89.6, 89.6], [89.6, 89.6, 89.6]),
    ([400, 400, 320], [1, 1, 1], None, None, [96, 96, 96], [96, 96, 96]),
    ([64, 48, 80], [1, 1, 1], None, None, [32, 32, 32], [16, 16, 16]),
    ([160, 128, 96], [0.8, 0.8, 1.25], None, None, [51.2, 51.2, 51.2], [25.6, 25.6, 25.6]),
    ([160, 128, 96], [1, 1, 1], [13, 20, 5], [120, 99, 70], [48, 48, 48], [24, 24, 24]),
    ([160, 128, 96], [1, 1, 1], [100, 60, 40], [160, 128, 96], [64, 64, 64], [64, 64, 64]),
    ([64, 64, 64], [1, 1, 1], [10, 10, 10], [20, 20, 20], [96, 96, 96], [96, 96, 96]),
    ([128, 128, 128], [1.5, 1.5, 1.5], None, None, [100, 100, 100], [33, 33, 33]),
]


def gen_grids():
    res = []
    for size, spacing, bs, be, psize, pstride in GRID_CASES:
        im = sitk.Image(size, sitk.sitkFloat32)
        im.SetSpacing(spacing)
        bs_ = copy.deepcopy(bs) if bs is not None else [0, 0, 0]
        be_ = copy.deepcopy(be) if be is not None else list(size)
        s, e = ref_it.image_partition_by_fixed_size(im, bs_, be_, copy.deepcopy(psize), copy.deepcopy(pstride), 16)
        res.append({'size': size, 'spacing': spacing, 'bbox_start': bs, 'bbox_end': be, 'partition_size': psize,
                    'partition_stride': pstride, 'n': len(s),
                    'starts': [[int(v) for v in p] for p in s], 'ends': [[int(v) for v in p] for p in e],
                    'bbox_start_after': [int(v) for v in bs_], 'bbox_end_after': [int(v) for v in be_]})
        print('grid', size, psize, pstride, len(s))
    json.dump(res, open(os.path.join(HERE, 'grids.json'), 'w'))


def gen_loss():
    out = {}
    g = torch.Generator().manual_seed(5)
    for c, shape in ((2, (2, 2, 8, 8, 8)), (5, (3, 5, 4, 8, 6))):
        logits = torch.randn(shape, generator=g) * 2
        probs = torch.softmax(logits, 1)
        # exact ties at 1/C on some voxels
        probs[:, :, 0, 0, :] = 1.0 / c
        target = torch.randint(0, c, (shape[0], 1) + shape[2:], generator=g).float()
        w = [1.0 + i for i in range(c)]
        p1 = probs.clone().requires_grad_(True)
        l1 = MultiDiceLoss(w, c, False)(p1, target)
        l1.backward()
        p2 = probs.clone().requires_grad_(True)
        l2 = FocalLoss(c, alpha=w, gamma=2, size_average=True, use_gpu=False)(p2, target)
        l2.backward()
        p3 = probs.clone().requires_grad_(True)
        l3 = FocalLoss(c, alpha=None, gamma=0, size_average=False, use_gpu=False)(p3, target)
        l3.backward()
        k = 'c%d_' % c
        out[k + 'probs'], out[k + 'target'], out[k + 'weights'] = probs.numpy(), target.numpy(), np.array(w, np.float32)
        out[k + 'dice'], out[k + 'dice_grad'] = l1.detach().numpy(), p1.grad.numpy()
        out[k + 'focal'], out[k + 'focal_grad'] = l2.detach().numpy(), p2.grad.numpy()
        out[k + 'focal_g0_sum'], out[k + 'focal_g0_sum_grad'] = l3.detach().numpy(), p3.grad.numpy()
        print('loss c=%d dice=%.6f focal=%.6f' % (c, float(l1), float(l2)))
    pb = torch.softmax(torch.randn((2, 2, 6, 6, 6), generator=g), 1)
    tb = torch.randint(0, 2, (2, 1, 6, 6, 6), generator=g).float()
    out['bin_probs'], out['bin_target'] = pb.numpy(), tb.numpy()
    out['bin_dice'] = BinaryDiceLoss()(pb.clone(), tb).numpy()
    np.savez_compressed(os.path.join(HERE, 'loss.npz'), **out)


SW_CASES = [
    # name, arch, cout, wseed, aseed, size xyz, psize, pstride, normalizer, vol seed
    ('sw_vnet_fixed', 'vnet', 2, 0, 3, [64, 48, 80], [32, 32, 32], [16, 16, 16], ('fixed', 50.0, 200.0, True), 11),
    ('sw_vnet_adaptive', 'vnet', 2, 1, None, [48, 48, 64], [32, 32, 32], [32, 32, 32], ('adaptive', 2.5), 12),
    ('sw_vbnet_fixed', 'vbnet', 5, 0, 4, [48, 64, 48], [32, 32, 32], [16, 32, 16], ('fixed', 0.0, 1.0, False), 13),
]


def synth_volume(seed, size_xyz, scale):
    x = seeded_input(seed, (1, 1, size_xyz[2], size_xyz[1], size_xyz[0]), 'smooth')[0, 0].numpy()
    return (x * scale).astype(np.float32)


# 64^3 patches (the reduced-precision bars are stated for full-size patches; the 32^3 cases above are too small for them).
# The probabilities of these larger volumes are committed on every second voxel per axis, the mask in full.
SW64_CASES = [
    ('sw64_vnet_overlap', 'vnet', 2, 0, None, [128, 128, 112], [64, 64, 64], [48, 48, 48], ('fixed', 0.0, 300.0, True), 21),
    ('sw64_vnet_tiled', 'vnet', 2, 2, None, [128, 64, 128], [64, 64, 64], [64, 64, 64], ('adaptive', 3.0), 22),
]


def gen_sliding_window(cases=None, fname='sliding_window.npz', sub=1):
    from easydict import EasyDict as edict
    out, meta = {}, []
    for name, arch, cout, wseed, aseed, size, psize, pstride, norm, vseed in (cases or SW_CASES):
        net = make_net(arch, 1, cout, wseed)
        if aseed is not None:
            net.load_state_dict(randomize_affine(net.state_dict(), aseed))
        model = edict()
        model.net = net
        model.spacing, model.max_stride, model.interpolation = [1.0, 1.0, 1.0], 16, 'LINEAR'
        model.in_channels, model.out_channels = 1, cout
        if norm[0] == 'fixed':
            model.crop_normalizers = [FixedNormalizer(norm[1], norm[2], norm[3])]
            scale = 300.0 if norm[2] > 1 else 1.0
        else:
            model.crop_normalizers = [AdaptiveNormalizer(norm[1])]
            scale = 300.0
        cfg = edict()
        cfg.partition_type, cfg.partition_size, cfg.partition_stride = 'SIZE', psize, pstride
        cfg.cpu_model_spacing_increase_ratio, cfg.cpu_partition_decrease_ratio = 1.0, 1.0
        cfg.pick_largest_cc, cfg.remove_small_cc = False, 0
        vol = synth_volume(vseed, size, scale)
        image = sitk.GetImageFromArray(vol)
        mean_probs, mask = ref_infer.segmentation_volume(model, cfg, image, None, None, False)
        probs = np.stack([sitk.GetArrayFromImage(p) for p in mean_probs], 0).astype(np.float32)
        out[name + '_probs'] = probs[:, ::sub, ::sub, ::sub].copy()
        out[name + '_mask'] = sitk.GetArrayFromImage(mask).astype(np.int8)
        meta.append([name, arch, cout, wseed, aseed, size, psize, pstride, list(norm), vseed, scale])
        print(name, probs.shape, 'fg frac', float((out[name + '_mask'] > 0).mean()))
    out['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, fname), **out)


def gen_cascade():
    """core/seg_infer.py:249-339 called the way the coarse->fine cascade calls it (:428-444): a bounding box restricts
    the patch grid.  Voxels outside the (max_stride-rounded) box are never visited: their overlap count is 0, the
    reference multiplies 0 by 1/0 = inf there, so their probabilities are NaN and their label is 0."""
    from easydict import EasyDict as edict
    net = make_net('vnet', 1, 2, 2)
    net.load_state_dict(randomize_affine(net.state_dict(), 6))
    model = edict()
    model.net = net
    model.spacing, model.max_stride, model.interpolation = [1.0, 1.0, 1.0], 16, 'LINEAR'
    model.in_channels, model.out_channels = 1, 2
    model.crop_normalizers = [FixedNormalizer(20.0, 250.0, True)]
    cfg = edict()
    cfg.partition_type, cfg.partition_size, cfg.partition_stride = 'SIZE', [32, 32, 32], [16, 16, 16]
    cfg.cpu_model_spacing_increase_ratio, cfg.cpu_partition_decrease_ratio = 1.0, 1.0
    cfg.pick_largest_cc, cfg.remove_small_cc = False, 0
    size, bs, be, vseed, scale = [64, 48, 64], [9, 5, 14], [49, 40, 50], 17, 300.0
    vol = synth_volume(vseed, size, scale)
    image = sitk.GetImageFromArray(vol)
    with np.errstate(all='ignore'):
        mean_probs, mask = ref_infer.segmentation_volume(model, cfg, image, list(bs), list(be), False)
    probs = np.stack([sitk.GetArrayFromImage(p) for p in mean_probs], 0).astype(np.float32)
    m = sitk.GetArrayFromImage(mask).astype(np.int8)
    meta = {'arch': 'vnet', 'cout': 2, 'wseed': 2, 'aseed': 6, 'size': size, 'psize': [32, 32, 32], 'pstride': [16, 16, 16],
            'norm': ['fixed', 20.0, 250.0, True], 'vseed': vseed, 'scale': scale, 'bbox_start': bs, 'bbox_end': be}
    print('cascade: visited fraction', float(np.isfinite(probs[0]).mean()), 'fg frac', float((m > 0).mean()))
    np.savez_compressed(os.path.join(HERE, 'cascade.npz'), probs=probs, mask=m, meta=np.array(json.dumps(meta)))


def gen_loss_ce():
    """loss/cross_entropy_loss.py:5-18 (loss.name = 'CE', core/seg_train.py:98-99): the reference's wrapper applied to
    the probabilities and targets of loss.npz - default construction, class weights, an ignored label, sum / none."""
    from segmentation3d.loss.cross_entropy_loss import CrossEntropyLoss as RefCE
    z = np.load(os.path.join(HERE, 'loss.npz'))
    out = {}
    for c in (2, 5):
        k = 'c%d_' % c
        probs, target = torch.from_numpy(z[k + 'probs']), torch.from_numpy(z[k + 'target'])
        w = torch.tensor([1.0 + 0.5 * i for i in range(c)])
        for tag, kw in (('ce', {}), ('ce_w', {'weight': w}), ('ce_ign', {'weight': w, 'ignore_index': 1}),
                        ('ce_sum', {'reduction': 'sum'}), ('ce_none', {'weight': w, 'ignore_index': 0, 'reduction': 'none'})):
            p = probs.clone().requires_grad_(True)
            l = RefCE(**kw)(p, target)
            (l if l.dim() == 0 else (l * torch.arange(l.numel(), dtype=torch.float32).view_as(l) / l.numel()).sum()).backward()
            out[k + tag], out[k + tag + '_grad'] = l.detach().numpy(), p.grad.numpy()
        out[k + 'weight'] = w.numpy()
        print('ce c=%d' % c, float(out[k + 'ce']), float(out[k + 'ce_w']), float(out[k + 'ce_ign']), float(out[k + 'ce_sum']))
    np.savez_compressed(os.path.join(HERE, 'loss_ce.npz'), **out)


def dataset_cases():
    """Three synthetic (image, mask) pairs with different sizes / spacings / origins; case 1 is thinner than the crop along
    z, case 2 has no voxel of label 2.  Duplicated in tests/test_host_logic.py (the reference package cannot be imported
    next to the product's)."""
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


def gen_dataset_sampling():
    """dataloader/dataset.py:140-209 with a recording stand-in for sitk.Resample: which crop (origin, spacing, size,
    interpolator) the UNMODIFIED reference asks for, per sampling method, under a seeded numpy RNG - plus the frame and
    case name it returns.  The resampling arithmetic itself (ITK) is not exercised."""
    import tempfile
    from segmentation3d.dataloader import dataset as ref_ds
    cases = dataset_cases()
    store, calls = {}, []
    tmp = tempfile.mkdtemp()
    lines = [str(len(cases))]
    for k, (im, lab, spacing, origin) in enumerate(cases):
        d = os.path.join(tmp, 'case%d' % k)
        os.makedirs(d)
        for name, arr in (('im.mha', im), ('seg.mha', lab)):
            path = os.path.join(d, name)
            open(path, 'w').close()
            img = sitk.GetImageFromArray(arr)
            img.SetSpacing(spacing), img.SetOrigin(origin)
            store[path] = img
            lines.append(path)
    with open(os.path.join(tmp, 'train.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')

    def read_image(path, pixel_id=None):
        src = store[path]
        return src._wrap(src._a.copy())

    def resample(image, size, transform, interp, origin, spacing, direction):
        calls.append({'size': [int(v) for v in size], 'origin': [float(v) for v in origin], 'spacing': [float(v) for v in spacing],
                      'interp': 'LINEAR' if interp == sitk.sitkLinear else 'NN'})
        out = sitk.Image([int(v) for v in size], sitk.sitkFloat32)
        out.SetOrigin(origin), out.SetSpacing(spacing), out.SetDirection(direction)
        return out

    sitk.ReadImage, sitk.Resample = read_image, resample
    sitk.Image.TransformIndexToPhysicalPoint = lambda self, idx: self.TransformContinuousIndexToPhysicalPoint(idx)
    out = {}
    for method in ('CENTER', 'GLOBAL', 'MASK', 'HYBRID'):
        ds = ref_ds.SegmentationDataset(os.path.join(tmp, 'train.txt'), num_classes=3, spacing=[1.0, 1.0, 1.2], crop_size=[32, 32, 32],
                                        sampling_method=method, random_translation=[5, 4, 3], random_scale=[0.9, 1.1],
                                        interpolation='LINEAR', crop_normalizers=[None])
        np.random.seed(1234)
        items = []
        for index in (0, 1, 2, 1, 0, 2, 2, 1):
            del calls[:]
            im_t, seg_t, frame, name = ds[index]
            assert tuple(im_t.shape) == (1, 32, 32, 32) and tuple(seg_t.shape) == (1, 32, 32, 32)
            items.append({'index': index, 'calls': [dict(c) for c in calls], 'frame': [float(v) for v in frame], 'name': name})
        out[method] = items
        print('dataset', method, items[0]['calls'][0]['origin'], items[0]['calls'][0]['spacing'])
    with open(os.path.join(HERE, 'dataset_sampling.json'), 'w') as f:
        json.dump(out, f)


def gen_samplers():
    """dataloader/sampler.py:6-53: the index streams of the unmodified single-process samplers under python's `random`."""
    import random
    from segmentation3d.dataloader import sampler as ref_sampler
    out = {}
    random.seed(5)
    out['concat_n7_e3_seed5'] = list(ref_sampler.EpochConcateSampler(list(range(7)), 3))
    out['resume_n5_e3_from2'] = list(ref_sampler.EpochConcateSamplerResume(list(range(5)), 3, 2))
    with open(os.path.join(HERE, 'samplers.json'), 'w') as f:
        json.dump(out, f)
    print('samplers', out)


def gen_train_step():
    """core/seg_train.py:83,119-127 on one synthetic batch: Adam(lr=1e-4, betas=(0.9,0.999))."""
    out = {}
    for name, arch, cout, lossname in (('vnet_dice', 'vnet', 2, 'Dice'), ('vbnet_focal', 'vbnet', 5, 'Focal')):
        net = make_net(arch, 1, cout, 0).train()
        g = torch.Generator().manual_seed(21)
        crops = torch.randn((2, 1, 16, 16, 32), generator=g)
        masks = torch.randint(0, cout, (2, 1, 16, 16, 32), generator=g).float()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999))
        if lossname == 'Dice':
            loss_func = MultiDiceLoss([1.0] * cout, cout, False)
        else:
            loss_func = FocalLoss(cout, alpha=[1.0] * cout, gamma=2, use_gpu=False)
        losses = []
        for _ in range(2):
            opt.zero_grad()
            loss = loss_func(net(crops), masks)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        sd = net.state_dict()
        out[name + '_losses'] = np.array(losses, np.float64)
        for k in ('in_block.conv.weight', 'down_64.down_conv.weight', 'up_32.up_conv.weight', 'out_block.conv2.weight',
                  'out_block.gn1.weight', 'up_128.up_gn.bias'):
            out[name + '/' + k] = sd[k].detach().numpy()
        out[name + '_grad_in_block'] = net.in_block.conv.weight.grad.numpy()
        print(name, losses)
    np.savez_compressed(os.path.join(HERE, 'train_step.npz'), **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['schema', 'forward', 'grids', 'loss', 'sw', 'sw64', 'train', 'cascade', 'ce', 'dataset', 'samplers']
    if 'schema' in which: gen_schema_and_hashes()
    if 'forward' in which: gen_forward()
    if 'grids' in which: gen_grids()
    if 'loss' in which: gen_loss()
    if 'sw' in which: gen_sliding_window()
    if 'sw64' in which: gen_sliding_window(SW64_CASES, 'sliding_window_64.npz', 2)
    if 'train' in which: gen_train_step()
    if 'cascade' in which: gen_cascade()
    if 'ce' in which: gen_loss_ce()
    if 'dataset' in which: gen_dataset_sampling()
    if 'samplers' in which: gen_samplers()
