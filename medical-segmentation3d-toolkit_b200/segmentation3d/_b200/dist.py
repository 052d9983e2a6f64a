"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on the box, gloo in CPU tests).

The hot path shards naturally:
  * inference over a case list      -> cases dealt rank::world, no collective;
  * inference of one volume         -> patches dealt rank::world, one all-reduce(sum) of the fp32
                                       accumulators before the overlap-count normalisation;
  * training                        -> data parallel, one bucketed all-reduce(sum)/world of the gradients
                                       per step (GroupNorm is per sample, so no statistics are exchanged).
These helpers hold that logic so it can be exercised with world_size-2 gloo groups on CPU.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items, rank=None, world_size=None):
    """Round-robin deal: rank r gets items[r::world] (reference order preserved inside a shard)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return list(items)[rank::world_size]


def sum_accumulators(acc):
    """all-reduce(sum) of the per-class probability accumulators (in place); no-op for world 1."""
    _, w = world()
    if w > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def allreduce_mean_grads(params, bucket_bytes=64 << 20):
    """Average gradients over ranks with flat buckets (a few large all-reduces: NVSwitch cost is launch
    latency, not link count).  Matches the reference's single-process DataParallel step, where the loss
    is the mean over the global batch: mean over ranks of per-rank batch means (equal per-rank batches)."""
    _, w = world()
    if w == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0
    def flush():
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(w)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()


def broadcast_params(module, src=0):
    """Make every rank start from rank `src`'s weights (what DataParallel's replicate does each step)."""
    _, w = world()
    if w > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)
