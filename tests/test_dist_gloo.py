"""world_size-2 gloo tests of the multi-GPU plumbing (runs on CPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'medical-segmentation3d-toolkit_b200'))
    sys.path.insert(0, root)
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from segmentation3d._b200 import dist as D
    from segmentation3d._b200.sliding import axis_counts
    from segmentation3d.utils.image3d import Image3d
    from segmentation3d.utils.image_tools import image_partition_by_fixed_size
    # ---- patch-sharded accumulation of one volume: sum of per-rank partial accumulators == serial accumulate
    size = [64, 48, 80]
    im = Image3d(np.zeros((1, 1, 1), np.float32))
    im.GetSize = lambda: tuple(size)
    starts, ends = image_partition_by_fixed_size(im, [0, 0, 0], list(size), [32, 32, 32], [16, 16, 16], 16)
    g = torch.Generator().manual_seed(0)
    probs = torch.rand((len(starts), 2, 32, 32, 32), generator=g)       # identical on every rank
    acc = torch.zeros((2, size[2], size[1], size[0]))
    mine = D.shard(list(range(len(starts))))
    assert mine == list(range(len(starts)))[rank::world]
    for i in mine:
        s, e = starts[i], ends[i]
        acc[:, s[2]:e[2], s[1]:e[1], s[0]:e[0]] += probs[i]
    D.sum_accumulators(acc)
    full = torch.zeros_like(acc)
    for i, (s, e) in enumerate(zip(starts, ends)):
        full[:, s[2]:e[2], s[1]:e[1], s[0]:e[0]] += probs[i]
    ok_acc = bool((acc - full).abs().max() <= 1e-5)
    cx, cy, cz = axis_counts(size, starts, ends)      # the count uses the FULL grid on every rank
    ok_cnt = int(cx.min()) >= 1 and int(cz.max()) == 2
    # ---- case sharding: disjoint cover
    cases = ['case_%d' % i for i in range(7)]
    got = [None] * world
    dist.all_gather_object(got, D.shard(cases))
    ok_cases = sorted(sum(got, [])) == sorted(cases)
    # ---- data-parallel gradient averaging + initial broadcast
    torch.manual_seed(100 + rank)
    lin = torch.nn.Linear(5, 3)
    D.broadcast_params(lin)
    w0 = [None] * world
    dist.all_gather_object(w0, lin.weight.detach().clone())
    ok_bcast = bool(torch.equal(w0[0], w0[1]))
    x = torch.full((4, 5), float(rank + 1))
    lin(x).sum().backward()
    local = lin.weight.grad.clone()
    D.allreduce_mean_grads(list(lin.parameters()), bucket_bytes=16)     # tiny buckets: exercise the flush logic
    both = [None] * world
    dist.all_gather_object(both, local)
    ok_grad = bool(torch.allclose(lin.weight.grad, (both[0] + both[1]) / 2))
    out.put((rank, ok_acc, ok_cnt, ok_cases, ok_bcast, ok_grad))
    dist.destroy_process_group()


def test_two_rank_sharding_and_collectives():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r


def _labels_merge_worker(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    # a 2 x 2 x 2 grid of non-overlapping 4^3 patches dealt rank::world; labels of the patches a rank did not run are 0
    rng = np.random.default_rng(0)
    full = rng.integers(0, 5, size=(8, 8, 8)).astype(np.int8)
    starts = [[x, y, z] for x in (0, 4) for y in (0, 4) for z in (0, 4)]
    local = np.zeros_like(full)
    for s in starts[rank::world]:
        local[s[2]:s[2] + 4, s[1]:s[1] + 4, s[0]:s[0] + 4] = full[s[2]:s[2] + 4, s[1]:s[1] + 4, s[0]:s[0] + 4]
    t = torch.from_numpy(local)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, bool(np.array_equal(t.numpy(), full))))
    dist.destroy_process_group()


def test_patch_sharded_labels_merge_by_max():
    """gather='labels' (core/seg_infer.py::segmentation_volume_device): with non-overlapping patches every voxel is
    labelled by exactly one rank, the others hold 0 there, so a max all-reduce of the int8 masks is the full mask."""
    import multiprocessing as mp
    from segmentation3d.core.seg_infer import deal_patches
    starts = [[x, y, z] for x in (0, 4) for y in (0, 4) for z in (0, 4)]
    assert deal_patches(starts, [[s[0] + 4, s[1] + 4, s[2] + 4] for s in starts], 0, 2)[1]          # tiled: ranks touch disjoint voxels
    starts = [[x, 0, 0] for x in (0, 2, 4)]
    assert not deal_patches(starts, [[s[0] + 4, 4, 4] for s in starts], 0, 2)[1]    # stride 2 < size 4: one chain, cannot be dealt whole
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_labels_merge_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


class _Patch(object):
    """minimal stand-in for pytest's monkeypatch inside spawned workers"""

    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _sharded_engine_worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, 'medical-segmentation3d-toolkit_b200'), root, os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    import test_sliding_wiring as tsw
    from oracle import init as oinit
    from segmentation3d.core.seg_infer import segmentation_volume_device
    lib = tsw._install(_Patch(), [])
    sd = oinit.randomize_affine(oinit.init_state_dict('vnet', 1, 2, 3), 4)
    vol = (tsw._seeded(5, (1, 1, 32, 64, 32))[0, 0].numpy() * 100).astype(np.float32)
    nd = {'type': 0, 'mean': 0.0, 'stddev': 100.0, 'clip': False}
    out = {}
    for name, cfg, gather in (('tiled_labels', {'partition_type': 'SIZE', 'partition_size': [32, 32, 32], 'partition_stride': [32, 32, 32]}, 'labels'),
                              ('tiled_probs', {'partition_type': 'SIZE', 'partition_size': [32, 32, 32], 'partition_stride': [32, 32, 32]}, 'probs'),
                              ('overlap_probs', {'partition_type': 'SIZE', 'partition_size': [32, 32, 32], 'partition_stride': [16, 16, 16]}, 'probs')):
        model, _ = tsw._model(sd, 2, nd, lib, batch=1)
        single_acc, single_mask = segmentation_volume_device(model, cfg, torch.from_numpy(vol))
        model, plan = tsw._model(sd, 2, nd, lib, batch=1)          # one patch per forward on both sides: identical arithmetic
        acc, mask = segmentation_volume_device(model, cfg, torch.from_numpy(vol), shard=(rank, world), gather=gather)
        ok = bool(torch.equal(mask, single_mask))
        if gather == 'probs':
            ok = ok and float((acc - single_acc).abs().max()) <= 1e-6
        out[name] = (ok, sum(plan.forwards))
    # 'labels' exchange reduced to rank 0 only (mask_root): the root holds the single-process mask, the other rank its partial one
    cfg = {'partition_type': 'SIZE', 'partition_size': [32, 32, 32], 'partition_stride': [32, 32, 32]}
    model, _ = tsw._model(sd, 2, nd, lib, batch=1)
    _, single_mask = segmentation_volume_device(model, cfg, torch.from_numpy(vol))
    model, _ = tsw._model(sd, 2, nd, lib, batch=1)
    _, mask = segmentation_volume_device(model, cfg, torch.from_numpy(vol), shard=(rank, world), gather='labels', mask_root=0)
    out['labels_to_root'] = (bool(torch.equal(mask, single_mask)) if rank == 0 else bool((mask <= single_mask).all()), 1)
    q.put((rank, out))
    dist.destroy_process_group()


def test_patch_sharded_engine_under_gloo_equals_single_process():
    """segmentation_volume_device(shard=(rank, world)) with the kernels emulated (tests/test_sliding_wiring.py) and a real
    2-rank gloo group: the 'probs' exchange (all-reduce of the accumulators) and the 'labels' exchange (local arg-max + max
    all-reduce, non-overlapping patches) both return the single-process mask on every rank, each rank having run only its
    share of the patches."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_engine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert res[r]['tiled_labels'] == (True, 1) and res[r]['tiled_probs'] == (True, 1)         # 2 patches, one per rank
        assert res[r]['overlap_probs'][0] is True
        assert res[r]['labels_to_root'][0] is True
    assert res[0]['overlap_probs'][1] + res[1]['overlap_probs'][1] == 3                            # 3 overlapping patches in all


def _ddp_train_worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, 'medical-segmentation3d-toolkit_b200'), root, os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    import test_autograd_wiring as taw
    from oracle import init as oinit
    from oracle import loss as oloss
    from oracle import net as onet
    from segmentation3d.core.seg_train import make_optimizer, train_step
    from segmentation3d.loss.multi_dice_loss import MultiDiceLoss
    from segmentation3d.network import vnet
    os.environ['SEG3D_ALLREDUCE_BUCKET_MB'] = '1'                    # several overlapped all-reduces per backward pass
    taw._install(_Patch(), [])
    sd = oinit.randomize_affine(oinit.init_state_dict('vnet', 1, 2, 0), 5)
    g = torch.Generator().manual_seed(21)
    crops = torch.randn((2, 1, 16, 16, 32), generator=g)
    masks = torch.randint(0, 2, (2, 1, 16, 16, 32), generator=g).float()
    # oracle: the reference's single-process step on the global batch of two crops
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_loss = oloss.multi_dice_loss(onet.forward_with_grad(params, crops), masks, [1.0, 1.0])
    ref_loss.backward()
    net = vnet.SegmentationNet(1, 2)
    net.load_state_dict(sd)
    net.b200_mode = 'fp32'
    net.train()
    opt = make_optimizer(net, 1e-4)
    lf = MultiDiceLoss([1.0, 1.0], 2, False)
    loss = train_step(net, opt, lf, crops[rank:rank + 1], masks[rank:rank + 1])        # this rank's crop only
    reduced = bool(getattr(net._plan, 'grads_reduced_in_backward', False))
    worst = 0.0
    for name, p in net.named_parameters():
        gr = params[name].grad
        worst = max(worst, float((p.grad - gr).abs().max()) / (float(gr.abs().max()) + 1e-12))
    q.put((rank, worst, reduced, float(loss.detach())))
    dist.destroy_process_group()


def test_data_parallel_step_under_gloo_equals_the_global_batch_gradient():
    """core/seg_train.py::train_step on two gloo ranks (one crop each, kernels emulated as in tests/test_autograd_wiring.py):
    the gradient all-reduce issued from inside the backward pass leaves, on every rank, the gradient of the reference's
    single-process step on the global batch."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, worst, reduced, loss in res:
        assert reduced, 'the overlapped all-reduce did not run'
        assert worst <= 1e-2, (rank, worst)
