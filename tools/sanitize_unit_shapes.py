"""Unit-shape pass over every kernel family of libseg3d_b200.so for compute-sanitizer (sanitizers row).

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_unit_shapes.py [infer] [train] [post]

Small shapes so that the instrumented run ends within a couple of minutes: sliding-window inference of a 48^3 volume
with overlapping 32^3 patches (gather, network forward in fp16 / fp32x / fp32, blend, finalize + argmax), VBNet C=5,
one bf16 training step of 2 x 32^3 crops (backward kernels, weight gradients, Dice), connected components and resampling.
Prints one line per section; the sanitizer's own summary says whether any access was out of bounds.
(compute-sanitizer is closed on the round-1 GPU pool - the call is refused - so the script was only run plain there,
as a quick pass over every kernel family.)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200'))
sys.path.insert(0, ROOT)
import torch                                                                   # noqa: E402


def make_net(arch, classes, mode):
    import importlib
    mod = importlib.import_module('segmentation3d.network.' + arch)
    torch.manual_seed(0)
    net = mod.SegmentationNet(1, classes)
    mod.parameters_kaiming_init(net)
    net.b200_mode = mode
    return net


def infer():
    from segmentation3d.core.seg_infer import make_model, segmentation_volume_device, segmentation_volume_host
    g = torch.Generator().manual_seed(3)
    vol = (torch.randn((48, 48, 48), generator=g) * 300.0).cuda()
    cfg = {'partition_type': 'SIZE', 'partition_size': [32] * 3, 'partition_stride': [16] * 3}
    for arch, classes, mode, norm in (('vnet', 2, 'fp16', {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True}),
                                      ('vbnet', 5, 'fp16', {'type': 1, 'clip_sigma': 3}),
                                      ('vnet', 2, 'fp32x', {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True}),
                                      ('vnet', 2, 'bf16', {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True}),
                                      ('vnet', 2, 'fp32', {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True})):
        net = make_net(arch, classes, mode).cuda().eval()
        model = make_model(net, spacing=[1.0, 1.0, 1.0], normalizer=norm)
        acc, mask = segmentation_volume_device(model, cfg, vol, batch=3)
        torch.cuda.synchronize()
        s = float(acc.sum(0).mean())
        print('infer %-5s C=%d %-5s sum_c p = %.6f, labels %s' % (arch, classes, mode, s, torch.unique(mask).tolist()), flush=True)
        assert abs(s - 1.0) < 1e-3
    host = torch.empty((48, 48, 48), dtype=torch.float32, pin_memory=True)
    host.copy_(vol)
    acc, hm = segmentation_volume_host(model, cfg, host, batch=3)
    torch.cuda.synchronize()
    print('infer host path: mask sum %d' % int(hm.sum()), flush=True)


def train():
    from segmentation3d.core.seg_train import make_optimizer, train_step
    from segmentation3d.loss.focal_loss import FocalLoss
    from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
    for arch, classes, mode, lf in (('vnet', 2, 'bf16', MultiDiceLoss([0.5, 0.5], 2, True)),
                                    ('vbnet', 3, 'bf16', FocalLoss(3, alpha=[1.0] * 3, gamma=2, use_gpu=True)),
                                    ('vnet', 2, 'fp32', MultiDiceLoss([0.5, 0.5], 2, True))):
        net = make_net(arch, classes, mode).cuda().train()
        opt = make_optimizer(net, 1e-4)
        g = torch.Generator().manual_seed(5)
        crops = torch.randn((2, 1, 32, 32, 32), generator=g).cuda()
        masks = torch.randint(0, classes, (2, 1, 32, 32, 32), generator=g).float().cuda()
        for _ in range(2):
            loss = train_step(net, opt, lf, crops, masks)
        torch.cuda.synchronize()
        print('train %-5s C=%d %-4s loss %.6f' % (arch, classes, mode, float(loss)), flush=True)


def post():
    from segmentation3d.utils.image3d import Image3d
    from segmentation3d.utils.image_tools import pick_largest_connected_component, remove_small_connected_component, resample_spacing
    g = torch.Generator().manual_seed(7)
    m = (torch.rand((24, 40, 56), generator=g) > 0.7).to(torch.int8).cuda()
    a = pick_largest_connected_component(Image3d(m), [1])
    b = remove_small_connected_component(Image3d(m), [1], 5)
    im = Image3d(torch.randn((20, 36, 44), generator=g).cuda(), spacing=(0.7, 0.9, 1.3))
    r = resample_spacing(im, [1.0, 1.0, 1.0], 16, 'LINEAR')
    torch.cuda.synchronize()
    print('post: largest cc %d voxels, >=5 %d voxels, resampled size %s' % (int(a.data.sum()), int(b.data.sum()), r.GetSize()), flush=True)


if __name__ == '__main__':
    todo = sys.argv[1:] or ['infer', 'train', 'post']
    for name in todo:
        {'infer': infer, 'train': train, 'post': post}[name]()
    print('sanitize_unit_shapes: done', flush=True)
