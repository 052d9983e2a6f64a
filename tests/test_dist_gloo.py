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
