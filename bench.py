"""Headline benchmark: sliding-window VNet inference on a synthetic 512x512x400 CT volume.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the hot path over one volume: crop+normalise 180 patches of 96^3
(partition_size = partition_stride = 96 mm at 1 mm spacing), VNet forward, overlap blend,
normalise by the overlap count, first-argmax mask.  Metric (BASELINE.json): Mvoxels/s =
volume voxels / time.  `value` is measured with the volume already in HBM; `e2e` goes through
the reference-facing call (segmentation3d.core.seg_infer.segmentation_volume) with a HOST volume
(pinned H2D copy in, int8 mask D2H out) inside the timed region.  With N > 1 every rank segments
its own volume (case-sharded batch inference, BASELINE config 5 style: no data-path collective,
weak scaling); `--shard patches` instead deals the patches of ONE volume r::N and sums the
accumulators with an NCCL all-reduce (strong scaling).

`--impl reference` times the reference's own CPU path (the oracle port of
core/seg_infer.segmentation_volume: two forwards per patch, numpy blend with the reference's
whole-volume copies) on the host cores, on a bounded sample of patches of the same workload.
`--task train` is the secondary metric (BASELINE configs[2], patches/s); its line carries a
`cpu_baseline` too (one oracle training step on one crop), and `--impl reference --task train`
prints that CPU arm on its own.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'medical-segmentation3d-toolkit_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = 'sliding-window infer Mvoxels/s (512x512x400 CT, VNet, 96^3 patches)'  # BASELINE.json headline (configs[1]); main() renames it for --arch vbnet
NORMALIZER = {'type': 0, 'mean': 0.0, 'stddev': 1000.0, 'clip': True}


def synth_ct(size_xyz, seed, device):
    """Seeded CT-like volume [z,y,x] float32 in HU: smooth low-frequency field + blobs + noise."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    X, Y, Z = size_xyz
    lo = torch.randn((1, 1, max(2, Z // 32), max(2, Y // 32), max(2, X // 32)), generator=g)
    lo = lo.to(device)
    field = torch.nn.functional.interpolate(lo, size=(Z, Y, X), mode='trilinear', align_corners=False)[0, 0]
    gd = torch.Generator(device=device).manual_seed(seed + 1) if device != 'cpu' else g
    noise = torch.randn((Z, Y, X), generator=gd, device=device)
    vol = (field * 600.0 + noise * 60.0 - 200.0).clamp_(-1000.0, 2000.0)
    return vol.float().contiguous()


def make_net(mode, arch='vnet', classes=2):
    import importlib
    mod = importlib.import_module('segmentation3d.network.' + arch)
    torch.manual_seed(0)
    net = mod.SegmentationNet(1, classes)
    mod.parameters_kaiming_init(net)
    net.b200_mode = mode
    return net


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        # samples under load = upper half of the clock samples
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {'sm_mhz': med, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        d = json.load(open(path))
        return d.get('hbm_gbs', 6536.0), d.get('bf16_tflops', 1641.1), d.get('bf16_tflops_sustained', 1382.6), 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'


def cpu_reference_sample(size_xyz, patch, stride, n_patches, seed):
    """Reference CPU path (oracle port) on the first n_patches of the workload. Returns (Mvox/s, seconds, cores)."""
    from oracle import init as oinit
    from oracle import sliding_window as osw
    sd = oinit.init_state_dict('vnet', 1, 2, 0)
    vol = synth_ct(size_xyz, seed, 'cpu').numpy()
    total = len(osw.partition_grid(size_xyz, [1, 1, 1], [0, 0, 0], list(size_xyz), [patch] * 3, [stride] * 3, 16)[0])
    t0 = time.time()
    osw.segmentation_volume(sd, vol, [1.0, 1.0, 1.0], NORMALIZER, 'SIZE', [patch] * 3, [stride] * 3, 16,
                            double_forward=True, faithful_copies=True, max_patches=n_patches)
    dt = time.time() - t0
    vox = float(size_xyz[0]) * size_xyz[1] * size_xyz[2] * n_patches / total
    return vox / dt / 1e6, dt, torch.get_num_threads(), total


def cpu_train_sample(patch, batch, steps=1):
    """Reference CPU training step (core/seg_train.py:119-127 through the oracle's autograd program: forward, Dice loss,
    backward, Adam) on `batch` crops of patch^3.  Returns (patches/s, seconds, cores)."""
    from oracle import init as oinit
    from oracle import loss as oloss
    from oracle import net as onet
    sd = oinit.init_state_dict('vnet', 1, 2, 0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(0)
    crops = torch.randn((batch, 1, patch, patch, patch), generator=g)
    masks = torch.randint(0, 2, (batch, 1, patch, patch, patch), generator=g).float()
    t0 = time.time()
    for _ in range(steps):
        opt.zero_grad()
        loss = oloss.multi_dice_loss(onet.forward_with_grad(params, crops), masks, [0.5, 0.5])
        loss.backward()
        opt.step()
    dt = time.time() - t0
    return batch * steps / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.task == 'train':
        vals = [cpu_train_sample(args.patch, 1) for _ in range(args.warmup + args.steps)][args.warmup:]
        value, dt, cores = float(np.mean([v[0] for v in vals])), float(np.mean([v[1] for v in vals])), vals[0][2]
        sample = '1 training step on 1 crop of %d^3 per timed step (forward, Dice, backward, Adam; fp32)' % args.patch
        print(json.dumps({
            'impl': 'reference', 'metric': 'train patches/s (VNet, 96^3 patches, batch %d/GPU, Dice, Adam)' % args.train_batch,
            'value': value, 'unit': 'patches/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'VNet(1,2) training step, crops [%d,1,%d^3] per GPU, MultiDiceLoss, Adam lr 1e-4 (BASELINE configs[2])'
                                   % (args.train_batch, args.patch), 'mode': args.mode, 'parallelism': 'dp%d' % args.gpus},
            'cpu_baseline': {'value': value, 'unit': 'patches/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))
        return
    size = [int(v) for v in args.volume.split(',')]
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, cores, total = cpu_reference_sample(size, args.patch, args.stride, args.ref_patches, 1234)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals])) * 1e3
    sample = '%d of %d patches (x2 forwards each, as core/seg_infer.py:230-234) through the reference loop incl. whole-volume numpy copies; linear in patch count' % (args.ref_patches, total)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Mvoxels/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, size),
        'cpu_baseline': {'value': value, 'unit': 'Mvoxels/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'Mvoxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args, size):
    return {'workload': '%s(1,%d) random-init sliding-window inference, volume %dx%dx%d @1mm, partition_size=%d mm, '
                        'partition_stride=%d mm (BASELINE configs[%d])' % ({'vnet': 'VNet', 'vbnet': 'VBNet'}[args.arch], args.classes,
                                                                           size[0], size[1], size[2], args.patch, args.stride,
                                                                           1 if args.arch == 'vnet' else 3),
            'patch_batch': args.batch, 'mode': args.mode, 'shard': args.shard, 'gather': args.gather if args.shard == 'patches' else None,
            'l2_policy': 'inputs larger than L2 (volume 419 MB + accumulators 839 MB per step)'}


def run_train(args):
    """BASELINE configs[2]: VNet, 96^3 patches, batch 8/GPU, MultiDiceLoss, Adam(lr 1e-4), data parallel."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(dev))
    from segmentation3d._b200 import dist as D
    from segmentation3d.core.seg_train import make_optimizer, train_step
    from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
    net = make_net(args.mode).to(dev).train()
    D.broadcast_params(net)
    opt = make_optimizer(net, 1e-4, (0.9, 0.999))
    lf = MultiDiceLoss([0.5, 0.5], 2, True)
    B, P = args.train_batch, args.patch
    g = torch.Generator(device=dev).manual_seed(rank)
    crops = torch.randn((B, 1, P, P, P), generator=g, device=dev)
    masks = torch.randint(0, 2, (B, 1, P, P, P), generator=g, device=dev).float()
    params = list(net.parameters())
    for _ in range(max(args.warmup, 3)):
        loss = train_step(net, opt, lf, crops, masks, params)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = train_step(net, opt, lf, crops, masks, params)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores = cpu_train_sample(P, 1)
        cpu = {'value': v, 'unit': 'patches/s', 'cores': cores, 'kind': 'port',
               'sample': '1 training step on 1 crop of %d^3 (forward, Dice, backward, Adam; fp32), %.1f s' % (P, dt)}
    if rank == 0:
        print(json.dumps({
            'cpu_baseline': cpu,
            'metric': 'train patches/s (VNet, 96^3 patches, batch %d/GPU, Dice, Adam)' % B, 'value': B * world / (ms * 1e-3),
            'unit': 'patches/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp16': 'f16', 'bf16': 'bf16', 'fp32': 'f32', 'fp32x': 'f16x2'}[args.mode], 'data': 'synthetic',
            'config': {'workload': 'VNet(1,2) training step, crops [%d,1,%d^3] per GPU, MultiDiceLoss, Adam lr 1e-4 (BASELINE configs[2])' % (B, P),
                       'mode': args.mode, 'parallelism': 'dp%d' % world}, 'loss': float(loss.item())}))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='fp16', choices=['fp16', 'bf16', 'fp32', 'fp32x'],
                    help="fp32x = strict parity on the tensor cores (f16 hi/lo split operands, fp32 accumulate)")
    ap.add_argument('--batch', type=int, default=20, help='patches per network forward')
    ap.add_argument('--volume', default='512,512,400')
    ap.add_argument('--patch', type=int, default=96)
    ap.add_argument('--stride', type=int, default=96)
    ap.add_argument('--shard', default='cases', choices=['cases', 'patches'])
    ap.add_argument('--gather', default='mask', choices=['mask', 'probs', 'labels'],
                    help="--shard patches: 'mask' = reduce-scatter + slab finalize + all-gather of the int8 mask; 'probs' = all-reduce; "
                         "'labels' = local arg-max + max all-reduce of the int8 mask when patches do not overlap (not yet verified on GPUs)")
    ap.add_argument('--ref-patches', type=int, default=4, help='patches in the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--layers', action='store_true', help='print the per-kernel roofline table to stderr')
    ap.add_argument('--task', default='infer', choices=['infer', 'train'],
                    help="'train': secondary metric, VNet 96^3 training step (BASELINE configs[2]) in patches/s")
    ap.add_argument('--train-batch', type=int, default=8)
    ap.add_argument('--arch', default='vnet', choices=['vnet', 'vbnet'], help='vbnet + --classes 5 = BASELINE configs[3]')
    ap.add_argument('--classes', type=int, default=2)
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.task == 'train':
        return run_train(args)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (the product has no CPU fallback)'
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(dev))

    from segmentation3d._b200 import lib
    from segmentation3d.core.seg_infer import segmentation_volume_device, segmentation_volume_host, make_model
    lib.load()
    size = [int(v) for v in args.volume.split(',')]
    nvox = float(size[0]) * size[1] * size[2]
    net = make_net(args.mode, args.arch, args.classes).to(dev).eval()
    model = make_model(net, spacing=[1.0, 1.0, 1.0], normalizer=NORMALIZER)
    cfg = {'partition_type': 'SIZE', 'partition_size': [args.patch] * 3, 'partition_stride': [args.stride] * 3}
    vol = synth_ct(size, 1234 + (rank if args.shard == 'cases' else 0), dev)
    host_vol = torch.empty(vol.shape, dtype=torch.float32, pin_memory=True)
    host_vol.copy_(vol)
    host_mask = torch.empty(vol.shape, dtype=torch.int8, pin_memory=True)
    shard = (rank, world) if (args.shard == 'patches' and world > 1) else None

    def step_device():
        return segmentation_volume_device(model, cfg, vol, batch=args.batch, shard=shard, gather=args.gather)

    def step_host():
        return segmentation_volume_host(model, cfg, host_vol, host_mask, batch=args.batch, shard=shard, gather=args.gather)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model['engine'].kernel_launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = model['engine'].kernel_launches - l0
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(step_device, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_host, max(2, args.steps // 2), 3)

    units = nvox * (world if args.shard == 'cases' else 1)
    value = units / (ms_dev * 1e-3) / 1e6
    e2e = units / (ms_e2e * 1e-3) / 1e6

    if rank == 0:
        hbm, tf_burst, tf_sust, which = peaks()
        # per-kernel table from one instrumented forward of a full patch batch
        plan = net._current_plan()
        ws, ops = plan.plan(args.batch, args.patch, args.patch, args.patch)
        plan.run_profiled(ws, ops)
        prof = plan.run_profiled(ws, ops)
        kinds = {}
        for m, ms in prof:
            k = kinds.setdefault(m['kind'], {'ms': 0.0, 'flops': 0.0, 'bytes': 0.0, 'n': 0})
            k['ms'] += ms; k['flops'] += m['flops']; k['bytes'] += m['bytes']; k['n'] += 1
        tot_ms = sum(k['ms'] for k in kinds.values())
        if args.layers:
            for m, ms in prof:
                sys.stderr.write('%-34s %-14s %8.3f ms %8.1f TFLOP/s %8.1f GB/s\n' % (
                    m['name'], m['kind'], ms, m['flops'] / ms / 1e9, m['bytes'] / ms / 1e6))
            for kname, k in sorted(kinds.items(), key=lambda kv: -kv[1]['ms']):
                sys.stderr.write('KIND %-14s n=%3d %8.3f ms (%4.1f%%) %8.1f TFLOP/s %8.1f GB/s\n' % (
                    kname, k['n'], k['ms'], 100 * k['ms'] / tot_ms, k['flops'] / k['ms'] / 1e9, k['bytes'] / k['ms'] / 1e6))
        top = max(kinds.items(), key=lambda kv: kv[1]['ms'])
        tk = top[1]
        if top[0].startswith('conv_tc'):
            roof = {'bound': 'tensor', 'achieved': tk['flops'] / tk['ms'] / 1e9, 'peak': tf_sust, 'unit': 'TFLOP/s'}
        else:
            roof = {'bound': 'hbm', 'achieved': tk['bytes'] / tk['ms'] / 1e6, 'peak': hbm, 'unit': 'GB/s'}
        roof['frac'] = roof['achieved'] / roof['peak']
        # DRAM bytes per launch of this kernel class from the committed ncu --set full capture (tools/make_profiles.py),
        # scaled to this run's patch batch; null when no capture of this class exists
        roof['traffic'] = None
        tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
        if os.path.isfile(tpath) and args.arch == 'vnet' and args.mode == 'fp16':
            tj = json.load(open(tpath))
            if top[0] in tj:
                roof['traffic'] = tj[top[0]]['dram_bytes_per_launch'] * args.batch / float(tj.get('_batch', args.batch))
                roof['traffic_source'] = tj.get('_source')
                roof['algorithmic_bytes_per_launch'] = tk['bytes'] / tk['n']
        roof['kernel'] = top[0]
        roof['share_of_forward'] = tk['ms'] / tot_ms
        roof['peak_source'] = which + ' (MEASURED_PEAKS.json, sustained)' if which == 'measured' else which
        roof['launches_per_forward'] = tk['n']
        roof['avg_launch_ms'] = tk['ms'] / tk['n']
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, dt, cores, total = cpu_reference_sample(size, args.patch, args.stride, args.ref_patches, 1234)
            cpu = {'value': v, 'unit': 'Mvoxels/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d of %d patches (x2 forwards, reference loop incl. whole-volume numpy copies), %.1f s' % (args.ref_patches, total, dt)}
        line = {
            'metric': METRIC if args.arch == 'vnet' else METRIC.replace('VNet', 'VBNet C=%d' % args.classes), 'value': value, 'unit': 'Mvoxels/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev, 'higher_is_better': True,
            'scaling': 'weak' if args.shard == 'cases' else 'strong', 'vs_baseline': None,
            'dtype': {'fp16': 'f16', 'bf16': 'bf16', 'fp32': 'f32', 'fp32x': 'f16x2'}[args.mode], 'data': 'synthetic',
            'config': workload_config(args, size),
            'e2e': {'value': e2e, 'unit': 'Mvoxels/s', 'h2d_bytes_per_step': int(nvox * 4), 'd2h_bytes_per_step': int(nvox),
                    'ms_per_step': ms_e2e},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
            'kernel_shares': {k: round(v['ms'] / tot_ms, 4) for k, v in kinds.items()},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
