"""Inference engine (drop-in for reference segmentation3d/core/seg_infer.py).

Same entry points, argument meaning and error behaviour as the reference:
  load_single_model / load_models      (:99-205)   model-folder + checkpoint layout unchanged
  segmentation_volume                  (:249-350)  resample -> grid -> patch loop -> blend -> argmax -> CC
  segmentation                         (:353-493)  txt / single file / folder input, same outputs on disk
  read_test_txt / read_test_csv / read_test_folder (:23-96)

What differs is where the work runs: the resampled volume, the per-class accumulators and the
mask stay on the GPU; patches are cropped/normalised, pushed through the network (several per
forward), blended and arg-maxed by the CUDA kernels of libseg3d_b200.so
(segmentation3d/_b200/sliding.py).  The reference's second forward per patch (:230-234) returns the
same bits as the first and is not repeated.  There is no CPU path: gpu_id < 0 raises.
"""
import copy
import glob
import importlib
import os
import time

import numpy as np
import torch

from segmentation3d._b200 import lib
from segmentation3d._b200.sliding import SlidingWindow, axis_counts
from segmentation3d.utils.attrdict import AttrDict as edict
from segmentation3d.utils.file_io import load_config, readlines
from segmentation3d.utils.image3d import (AsyncImageWriter, Image3d, as_image3d, io_threads, prefetch_images, read_image,  # noqa: F401
                                          write_image)
from segmentation3d.utils.image_tools import (get_bounding_box, image_partition_by_fixed_size, is_identity_resample,  # noqa: F401
                                              resample, resample_spacing)
from segmentation3d.utils.model_io import get_checkpoint_folder
from segmentation3d.utils.normalizer import normalizer_from_dict

DEFAULT_PATCH_BATCH = int(os.environ.get('SEG3D_PATCH_BATCH', '0'))      # 0: chosen from the patch size


def default_patch_batch(patch_voxels):
    """patches per network forward: as many as fit ~40 GB of fp16 workspace (~430 B/voxel), at most 36 (measured on B200,
    profiles/r02_batch_graph_sweep.txt: 20 -> 36..45 patches per forward is +2-4 %; 36 is one z layer of BASELINE configs[1]).
    GroupNorm is per sample, so the batch size changes throughput, not results."""
    if DEFAULT_PATCH_BATCH > 0:
        return DEFAULT_PATCH_BATCH
    return int(min(36, max(1, 40e9 // (430.0 * max(1, patch_voxels)))))


# ---- test-list readers -----------------------------------------------------------------------------
IMAGE_SUFFIXES = ('.mhd', '.nii', '.hdr', '.nii.gz', '.mha', '.image3d')      # search order of the reference (:81)


def read_test_txt(txt_file):
    """First line = number of cases, then exactly that many '<case_name> <path>' lines (:23-46); every path must exist."""
    lines = readlines(txt_file)
    case_num = int(lines[0])
    if len(lines) - 1 != case_num:
        raise ValueError('case num do not equal path num!')
    names, paths = [], []
    for line in lines[1:1 + case_num]:
        tokens = line.strip().split()
        if len(tokens) < 2:
            raise ValueError('invalid line: %s' % line)
        if not os.path.isfile(tokens[1]):
            raise ValueError('image not exist: {}'.format(tokens[1]))
        names.append(tokens[0])
        paths.append(tokens[1])
    return names, paths


def read_test_csv(csv_file, mode='test'):
    """csv with columns image_name,image_path[,mask_path] (:49-67)."""
    import pandas as pd
    df = pd.read_csv(csv_file)
    if mode == 'test':
        return df['image_name'].tolist(), df['image_path'].tolist()
    if mode in ('train', 'validation'):
        return df['image_path'].tolist(), df['mask_path'].tolist()
    raise ValueError('Unsupported mode type.')


def read_test_folder(folder_path, is_dicom_folder=False):
    """Every image file of the folder, all suffixes sorted TOGETHER by path; the case name is the file name cut at the first
    known suffix found in it ('case1.mha' -> 'case1', 'case2.nii.gz' -> 'case2'), as the reference does (:70-96), so results
    land in <out>/case1/.  A DICOM folder is one case named after the folder."""
    if is_dicom_folder:
        return [os.path.split(folder_path)[1]], [folder_path]
    found = []
    for suf in IMAGE_SUFFIXES:
        found.extend(glob.glob(os.path.join(folder_path, '*' + suf)))
    names, paths = [], []
    for path in sorted(found):
        name = os.path.basename(path)
        cuts = [name.find(suf) for suf in IMAGE_SUFFIXES]
        hit = next((c for c in cuts if c != -1), -1)
        names.append(name[:hit] if hit != -1 else name)
        paths.append(path)
    return names, paths


def shard_case_list(names, paths, rank, world):
    """Batch inference over a case list shards by case (BASELINE configs[4]): rank r of a `torch.distributed.run`
    launch segments cases r, r+world, ...; cases are independent, so there is no collective on the data path."""
    if world <= 1:
        return list(names), list(paths)
    if not 0 <= rank < world:
        raise ValueError('rank %d outside world of %d' % (rank, world))
    return list(names[rank::world]), list(paths[rank::world])


def launch_rank():
    """(rank, world, local_rank) of a one-process-per-GPU launch (torchrun environment), (0, 1, None) otherwise."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1, None
    return int(os.environ.get('RANK', '0')), world, int(os.environ.get('LOCAL_RANK', os.environ.get('RANK', '0')))


# ---- model loading -----------------------------------------------------------------------------------
def _device_for(gpu_id):
    if gpu_id is None or int(gpu_id) < 0:
        raise RuntimeError('segmentation3d (B200 build) has no CPU inference path: pass gpu_id >= 0')
    if not torch.cuda.is_available():
        raise RuntimeError('segmentation3d (B200 build): no CUDA device visible')
    return torch.device('cuda:%d' % int(gpu_id)) if int(gpu_id) < torch.cuda.device_count() else torch.device('cuda:0')


def make_model(net, spacing, normalizer=None, max_stride=16, interpolation='LINEAR', batch=None):
    """Wrap an on-device network into the model record `segmentation_volume` consumes."""
    model = edict()
    model.net = net
    model.spacing, model.max_stride, model.interpolation = list(spacing), max_stride, interpolation
    model.in_channels, model.out_channels = net.in_channels, net.out_channels
    if normalizer is None:
        model.crop_normalizers = None
    else:
        model.crop_normalizers = [normalizer_from_dict(normalizer) if isinstance(normalizer, dict) else normalizer]
    dict.__setitem__(model, 'engine', SlidingWindow(net._current_plan(), batch or 0))
    return model


def load_single_model(model_folder, gpu_id=0):
    """<model_folder>/checkpoints/chk_<latest>/params.pth -> model record on cuda:<gpu_id>."""
    assert os.path.isdir(model_folder), 'Model folder does not exist: {}'.format(model_folder)
    device = _device_for(gpu_id)
    chk_dir = get_checkpoint_folder(os.path.join(model_folder, 'checkpoints'), -1)
    state = torch.load(os.path.join(chk_dir, 'params.pth'), map_location='cpu', weights_only=False)
    net_module = importlib.import_module('segmentation3d.network.' + state['net'])
    net = net_module.SegmentationNet(state['in_channels'], state['out_channels'])
    sd = state['state_dict']
    if any(k.startswith('module.') for k in sd):          # written under nn.DataParallel (core/seg_train.py:77,148)
        sd = {k[7:]: v for k, v in sd.items()}
    net.load_state_dict(sd)
    net = net.to(device).eval()
    norms = state['crop_normalizers']
    for d in norms:
        if d['type'] not in (0, 1):
            raise ValueError('Unsupported normalization type.')
    model = make_model(net, state['spacing'], norms[0] if norms else None, state['max_stride'], state['interpolation'])
    model.crop_normalizers = [normalizer_from_dict(d) for d in norms]
    return model


def load_models(model_folder, gpu_id=0):
    assert os.path.isdir(model_folder), 'Model folder does not exist: {}'.format(model_folder)
    models = edict()
    dict.__setitem__(models, 'infer_cfg', load_config(os.path.join(model_folder, 'infer_config.py')))
    cfg = models.infer_cfg
    scale = cfg.general.single_scale
    if scale not in ('coarse', 'fine', 'DISABLE'):
        raise ValueError('Unsupported single scale type!')
    coarse = fine = None
    if scale in ('coarse', 'DISABLE'):
        coarse = load_single_model(os.path.join(model_folder, cfg.coarse.model_name), gpu_id)
    if scale in ('fine', 'DISABLE'):
        fine = load_single_model(os.path.join(model_folder, cfg.fine.model_name), gpu_id)
    dict.__setitem__(models, 'coarse_model', coarse)
    dict.__setitem__(models, 'fine_model', fine)
    return models


# ---- the hot path --------------------------------------------------------------------------------------
def _cfg_get(cfg, key, default=None):
    return cfg[key] if key in cfg else default


def _grid(model, cfg, size_xyz, spacing, bbox_start_voxel, bbox_end_voxel, use_gpu=True):
    ptype = cfg['partition_type']
    if ptype == 'DISABLE':
        return [[0, 0, 0]], [[int(v) for v in size_xyz]]
    if ptype != 'SIZE':
        raise ValueError('Unsupported partition type!')
    psize, pstride = copy.deepcopy(list(cfg['partition_size'])), copy.deepcopy(list(cfg['partition_stride']))
    if not use_gpu:
        r = _cfg_get(cfg, 'cpu_partition_decrease_ratio', 1.0)
        psize, pstride = [v * r for v in psize], [v * r for v in pstride]
    if bbox_start_voxel is not None and bbox_end_voxel is not None:
        bs = [max(0, int(v)) for v in bbox_start_voxel]
        be = [min(int(bbox_end_voxel[a]), int(size_xyz[a])) for a in range(3)]
    else:
        bs, be = [0, 0, 0], [int(v) for v in size_xyz]
    frame = Image3d(np.empty((0, 0, 0), np.float32), spacing)
    frame.GetSize = lambda: tuple(int(v) for v in size_xyz)
    return image_partition_by_fixed_size(frame, bs, be, psize, pstride, model['max_stride'])


def segmentation_voi(model, iso_image, start_voxel, end_voxel, use_gpu):
    """Probability maps of one volume of interest [start_voxel, end_voxel) of an image already at the model spacing
    (reference :208-246): crop -> crop normaliser -> network -> one image per class carrying the VOI's frame.  The
    reference averages two forwards of the same tensor (:230-234), which returns the bits of one.  The sliding-window
    engine does not come through here (it crops, normalises and blends whole patch batches on the device); this is the
    single-VOI call kept for callers of the reference API."""
    img = as_image3d(iso_image)
    s, e = [int(v) for v in start_voxel], [int(v) for v in end_voxel]
    data = img.data[s[2]:e[2], s[1]:e[1], s[0]:e[0]]
    roi = Image3d(data, img.GetSpacing(), img.TransformContinuousIndexToPhysicalPoint([float(v) for v in s]), img.GetDirection())
    if model['crop_normalizers'] is not None:
        roi = model['crop_normalizers'][0](roi)
    dev = next(model['net'].parameters()).device
    x = roi.data if torch.is_tensor(roi.data) else torch.from_numpy(np.ascontiguousarray(roi.data))
    x = x.to(device=dev, dtype=torch.float32).unsqueeze(0).unsqueeze(0)
    with torch.no_grad():
        probs = model['net'](x)
    assert model['out_channels'] == probs.shape[1]
    maps = []
    for idx in range(model['out_channels']):
        m = Image3d(probs[0][idx])
        m.CopyInformation(roi)
        maps.append(m)
    return maps


def overlap_components(starts, ends):
    """Groups of patches that overlap, directly or through a chain.  The reference grid (utils/image_tools.py:200-216) is the
    Cartesian product of per-axis box lists, so two patches overlap iff their boxes overlap on every axis: merging the
    overlapping intervals of each axis and taking the product of the merged groups gives the connected components.  With
    partition_stride >= partition_size only the LAST box of an axis, clamped back into the volume (:209-213), overlaps its
    neighbour, so the components hold 1, 2, 4 or 8 patches.  Returned in z-major order as lists of patch indices."""
    group_of = []
    for a in range(3):
        boxes = sorted(set((int(s[a]), int(e[a])) for s, e in zip(starts, ends)))
        gid, reach, ids = -1, None, {}
        for lo, hi in boxes:
            if reach is None or lo >= reach:
                gid, reach = gid + 1, hi
            else:
                reach = max(reach, hi)
            ids[lo] = gid
        group_of.append(ids)
    comps = {}
    for i, s in enumerate(starts):
        key = (group_of[2][int(s[2])], group_of[1][int(s[1])], group_of[0][int(s[0])])
        comps.setdefault(key, []).append(i)
    return [comps[k] for k in sorted(comps)]


def deal_patches(starts, ends, rank, world, max_imbalance=1.15):
    """Patches of ONE volume for rank `rank` of `world`.  Returns (patches, disjoint).
    disjoint = True: whole overlap components (see above) are dealt, in z-major order, to the rank whose share of the patch
    count their midpoint falls in.  No voxel is then touched by two ranks: a rank's accumulators are final where its patches
    lie and exactly zero elsewhere, so it can count-normalise and arg-max locally and the ranks exchange LABELS.  A rank's
    patches span one or two z layers of the lattice: it reads, accumulates and finishes only that z range of the volume.
    disjoint = False (a component is too large to balance, e.g. partition_stride < partition_size chains every patch to its
    neighbour): single patches in z-major order, cut into `world` runs whose lengths differ by at most one; the ranks must
    then exchange probability sums."""
    n = len(starts)
    comps = overlap_components(starts, ends)
    # Every rank owns the window [rank*n/world, (rank+1)*n/world) of the z-major patch order, so a z group of components
    # (one z layer of the lattice; the last two layers when the last box was clamped back) is shared by the few ranks whose
    # windows reach into it, each with the capacity of that overlap.  Inside a group the components go, largest first, to the
    # rank with the most capacity left: loads end within a patch or two of n/world (180 patches over 8 ranks: 22 or 23 each),
    # and a rank still touches only the z layers of its window.
    groups = {}
    for comp in comps:
        groups.setdefault(min(int(starts[i][2]) for i in comp), []).append(comp)
    loads, a, owner, sharers = [0] * world, 0, [], []
    for z in sorted(groups):
        gcomps = sorted(groups[z], key=lambda c: -len(c))
        b = a + sum(len(c) for c in gcomps)
        cap = {}
        for r in range(world):
            lo, hi = n * r / float(world), n * (r + 1) / float(world)
            ov = min(b, hi) - max(a, lo)
            if ov > 1e-9:
                cap[r] = ov
        for comp in gcomps:
            r = max(cap, key=lambda k: (cap[k], -k))
            cap[r] -= len(comp)
            loads[r] += len(comp)
            owner.append([comp, r])
            sharers.append(tuple(cap))
        a = b
    # refinement: while the fullest rank can hand one of its components to a rank of the same z group that stays below it
    for _ in range(4 * world):
        rmax = max(range(world), key=lambda k: loads[k])
        best = None
        for j, (comp, r) in enumerate(owner):
            if r != rmax:
                continue
            for r2 in sharers[j]:
                if r2 != rmax and loads[r2] + len(comp) < loads[rmax] and (best is None or len(comp) < len(owner[best[0]][0])):
                    best = (j, r2)
        if best is None:
            break
        j, r2 = best
        loads[rmax] -= len(owner[j][0])
        loads[r2] += len(owner[j][0])
        owner[j][1] = r2
    mine = [i for comp, r in owner if r == rank for i in comp]
    if max(loads) <= max_imbalance * -(-n // world):
        return [starts[i] for i in mine], True
    order = sorted(range(n), key=lambda i: (starts[i][2], starts[i][1], starts[i][0]))
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    return [starts[i] for i in order[lo:hi]], False


def shard_plan(model, cfg, shape_zyx, shard, bbox_start_voxel=None, bbox_end_voxel=None, use_gpu=True, spacing=None):
    """Grid of a volume [Z,Y,X] and this rank's part of it: (starts, ends, mine, (z_lo, z_hi), disjoint).  z range = the planes
    this rank's patches read and write ((0, Z) without sharding)."""
    Z, Y, X = [int(v) for v in shape_zyx]
    starts, ends = _grid(model, cfg, [X, Y, Z], spacing or model['spacing'], bbox_start_voxel, bbox_end_voxel, use_gpu)
    if shard is None or shard[1] <= 1:
        return starts, ends, starts, (0, Z), True
    mine, disjoint = deal_patches(starts, ends, shard[0], shard[1])
    pz = ends[0][2] - starts[0][2]
    if not mine:
        return starts, ends, mine, (0, 0), disjoint
    return starts, ends, mine, (min(s[2] for s in mine), max(s[2] for s in mine) + pz), disjoint


def segmentation_volume_device(model, cfg, vol, batch=None, shard=None, bbox_start_voxel=None, bbox_end_voxel=None,
                               use_gpu=True, spacing=None, z_ready=None, mask_sink=None, gather='probs', mask_root=None):
    """Device-resident core of segmentation_volume: `vol` is a CUDA float32 [z,y,x] tensor already at
    the model spacing.  Returns (mean_probs [C,z,y,x] fp32, mask [z,y,x] int8) on the device.
    shard=(rank, world): the patches of this ONE volume are dealt over the ranks of the default process group (NCCL) in
    contiguous z runs (`deal_patches`); every rank touches only the z range of its patches.  The exchange step:
      gather='labels' (taken when the patches can be dealt so that no voxel is touched by two ranks - `deal_patches`; true for
          partition_stride >= partition_size, the BASELINE configs): count-normalise + arg-max the local z range and merge the
          int8 masks with ONE max all-reduce (1 byte per voxel on the wire).  Returned probabilities: this rank's z range.
      gather='mask' (needs Z % world == 0; also what 'labels' falls back to otherwise): per-class reduce-scatter
          of z slabs of the fp32 accumulators, count normalisation + arg-max on the local slab, all-gather of the int8 mask;
          the returned probabilities are this rank's slab [C, Z/world, Y, X].
      gather='probs': all-reduce of the full maps; every rank returns the full probability maps.
    mask_sink=(host_mask, stream): with z_ready, every z slab is normalised, arg-maxed and copied to the (pinned) host
    mask on `stream` as soon as no remaining patch touches it."""
    eng = model['engine']
    eng.plan = model['net']._current_plan()
    Z, Y, X = vol.shape
    sharded = shard is not None and shard[1] > 1
    starts, ends, mine, (z_lo, z_hi), disjoint = shard_plan(model, cfg, vol.shape, shard, bbox_start_voxel, bbox_end_voxel, use_gpu, spacing)
    norm = model['crop_normalizers'][0].to_dict() if model['crop_normalizers'] else None
    patch = [ends[0][a] - starts[0][a] for a in range(3)]
    C = model['out_channels']
    counts = axis_counts([X, Y, Z], starts, ends)
    if bbox_start_voxel is not None:
        # voxels outside the partitioned box have count 0: the reference divides by zero there (inf*0 = nan
        # -> argmax 0); keep them at probability 0 / label 0 by treating the count as 1.
        counts = [np.maximum(c, 1) for c in counts]
    pv = patch[0] * patch[1] * patch[2]
    if batch:
        eng.batch = int(batch)
    elif eng.batch <= 0:
        eng.batch = default_patch_batch(pv)
    if sharded and gather == 'labels' and disjoint:
        import torch.distributed as dist
        # local z range only: accumulators, count normalisation and arg-max cover [z_lo, z_hi); the volume may be resident
        # for that range alone (segmentation_volume_host uploads nothing else)
        mask = torch.zeros((Z, Y, X), dtype=torch.int8, device=vol.device)
        acc = torch.zeros((C, z_hi - z_lo, Y, X), dtype=torch.float32, device=vol.device)
        if mine:
            local = sorted([[s[0], s[1], s[2] - z_lo] for s in mine], key=lambda s: s[2])
            keep = eng.batch
            budget = max(eng.batch, int(40e9 // (430.0 * pv)))
            if z_ready is None:
                # one forward over all of this rank's patches when they fit the workspace budget (a 12 + 11 split of 23 patches
                # runs two latency-bound half batches)
                chunks = [local] if len(local) <= budget else [local[i:i + eng.batch] for i in range(0, len(local), eng.batch)]
            else:
                # the z range is still being uploaded slab by slab (segmentation_volume_host): one forward per z layer of the
                # lattice (layers merged while they stay within a patch batch), each waiting only for the planes it reads, so
                # the upload of the later layers hides behind the first forward
                chunks = []
                for s in local:
                    if chunks and (s[2] == chunks[-1][-1][2] or len(chunks[-1]) + sum(1 for t in local if t[2] == s[2]) <= eng.batch) \
                            and len(chunks[-1]) < budget:
                        chunks[-1].append(s)
                    else:
                        chunks.append([s])
            cur = torch.cuda.current_stream() if z_ready is not None else None
            for chunk in chunks:
                if z_ready is not None:
                    zmax = z_lo + max(s[2] for s in chunk) + patch[2]
                    for z1, ev in z_ready:
                        cur.wait_event(ev)
                        if z1 >= zmax:
                            break
                eng.batch = max(keep, len(chunk)) if len(chunk) <= budget else keep
                eng.accumulate(vol[z_lo:z_hi], chunk, patch, norm, acc)
            eng.batch = keep
            eng.finalize(acc, [counts[0], counts[1], np.ascontiguousarray(counts[2][z_lo:z_hi])], mask=mask[z_lo:z_hi])
        if mask_root is None:
            dist.all_reduce(mask, op=dist.ReduceOp.MAX)
        else:                 # only one process consumes the merged mask (it writes the result): reduce instead of all-reduce
            dist.reduce(mask, dst=int(mask_root), op=dist.ReduceOp.MAX)
        return acc, mask
    acc = torch.zeros((C, Z, Y, X), dtype=torch.float32, device=vol.device)
    progressive = z_ready is not None and mask_sink is not None and not sharded
    mask = torch.empty((Z, Y, X), dtype=torch.int8, device=vol.device) if progressive else None
    if z_ready is None:
        eng.accumulate(vol, mine, patch, norm, acc)
    else:
        # the volume is still being uploaded slab by slab (segmentation_volume_host): run the patches in order of
        # their last z plane and make each batch wait only for the slabs it reads.  The blend is order-independent.
        mine = sorted(mine, key=lambda s: s[2])
        cur = torch.cuda.current_stream()
        z_done = 0
        for b0 in range(0, len(mine), eng.batch):
            chunk = mine[b0:b0 + eng.batch]
            zmax = max(s[2] for s in chunk) + patch[2]
            for z1, ev in z_ready:
                cur.wait_event(ev)
                if z1 >= zmax:
                    break
            eng.accumulate(vol, chunk, patch, norm, acc)
            if progressive:
                rest = mine[b0 + eng.batch:]
                z_final = min(s[2] for s in rest) if rest else Z     # patches are sorted by their first z plane
                if z_final > z_done:
                    eng.finalize(acc, counts, z_range=(z_done, z_final), mask=mask)
                    host_mask, side = mask_sink
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    side.wait_event(ev)
                    with torch.cuda.stream(side):
                        host_mask[z_done:z_final].copy_(mask[z_done:z_final], non_blocking=True)
                    z_done = z_final
        if progressive:
            cur.wait_stream(mask_sink[1])
            return acc, mask
    if sharded:
        import torch.distributed as dist
        r, w = shard
        if gather in ('mask', 'labels') and Z % w == 0:
            zs = Z // w
            slab = torch.empty((C, zs, Y, X), dtype=torch.float32, device=acc.device)
            # class c of the accumulators is [world][zs, Y, X] as it lies in memory: one reduce-scatter per class, no re-layout
            for c in range(C):
                dist.reduce_scatter_tensor(slab[c], acc[c], op=dist.ReduceOp.SUM)
            mask_slab = eng.finalize(slab, [counts[0], counts[1], np.ascontiguousarray(counts[2][r * zs:(r + 1) * zs])])
            mask = torch.empty((Z, Y, X), dtype=torch.int8, device=acc.device)
            dist.all_gather_into_tensor(mask, mask_slab)
            return slab, mask
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    mask = eng.finalize(acc, counts)
    return acc, mask


def segmentation_volume_host(model, cfg, host_vol, host_mask=None, batch=None, shard=None, gather='probs', mask_root=None):
    """End-to-end call with HOST buffers: (pinned) float32 [z,y,x] in, int8 mask out; probabilities stay
    on the device and are returned as a tensor.  Copies are issued on the current stream.  With shard=(rank, world) a
    rank uploads only the z planes its patches read (all of them for the overlapping-patch gathers).  mask_root = r (sharded
    'labels' exchange only): the merged mask is reduced to rank r alone and only that process copies it to its host buffer
    (the others return their partial device mask) - one 105 MB device-to-host copy per volume instead of one per rank."""
    dev = next(model['net'].parameters()).device
    vol = torch.empty(host_vol.shape, dtype=torch.float32, device=dev)
    z_ready = None
    Z = host_vol.shape[0]
    z_lo, z_hi = 0, Z
    if shard is not None and shard[1] > 1:
        _, _, _, (z_lo, z_hi), _ = shard_plan(model, cfg, host_vol.shape, shard)
    if host_vol.is_pinned() and cfg['partition_type'] == 'SIZE' and os.environ.get('SEG3D_OVERLAP_UPLOAD', '1') != '0':
        # upload z slabs on a side stream so the first patches start while the rest of the volume is in flight
        side = _side_stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        step = max(16, int(cfg['partition_size'][2] / float(model['spacing'][2]) + 0.5))
        z_ready = []
        with torch.cuda.stream(side):
            for z0 in range(z_lo, z_hi, step):
                z1 = min(z_hi, z0 + step)
                vol[z0:z1].copy_(host_vol[z0:z1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
                z_ready.append((z1, ev))
    else:
        vol[z_lo:z_hi].copy_(host_vol[z_lo:z_hi], non_blocking=True)
    if host_mask is None:
        host_mask = torch.empty(host_vol.shape, dtype=torch.int8, pin_memory=True)
    sharded = shard is not None and shard[1] > 1
    sink = (host_mask, _side_stream(dev)) if (z_ready is not None and host_mask.is_pinned() and not sharded) else None
    acc, mask = segmentation_volume_device(model, cfg, vol, batch=batch, shard=shard, z_ready=z_ready, mask_sink=sink, gather=gather,
                                           mask_root=mask_root if sharded else None)
    if sharded and mask_root is not None and gather == 'labels' and shard[0] != int(mask_root):
        return acc, mask
    if sink is None:
        host_mask.copy_(mask, non_blocking=True)
    return acc, host_mask


_SIDE_STREAMS = {}


def _side_stream(dev):
    key = str(dev)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _cc_filter_device(mask, labels, min_size=0):
    """pick_largest_connected_component (min_size == 0) / remove_small_connected_component (min_size = threshold)
    (utils/image_tools.py:380-432; 26-connectivity) on the device: int8 CUDA mask [z,y,x] -> filtered int8 mask."""
    if not (torch.is_tensor(mask) and mask.is_cuda):
        raise RuntimeError('seg3d_b200: connected-component filtering runs on CUDA masks only (no CPU fallback)')
    mask = mask.contiguous()
    Z, Y, X = mask.shape
    out = torch.zeros_like(mask)
    parent = torch.empty((Z * Y * X,), dtype=torch.int32, device=mask.device)
    size = torch.empty((Z * Y * X,), dtype=torch.int32, device=mask.device)
    best = torch.empty((1,), dtype=torch.int64, device=mask.device)
    with torch.cuda.device(mask.device):
        for lab in labels:
            lib.call('seg3d_cc_filter', lib.ptr(mask), Z, Y, X, int(lab), int(min_size), lib.ptr(parent), lib.ptr(size),
                     lib.ptr(best), lib.ptr(out), lib.stream_ptr())
    return out


def segmentation_volume(model, cfg, image, bbox_start_voxel, bbox_end_voxel, use_gpu):
    """Segment one volume.  Returns (mean_probs: list of per-class images, mask image) like the
    reference; the images are backed by CUDA tensors (`.to_numpy()` copies to the host on demand)."""
    image = as_image3d(image)
    model_spacing = list(copy.deepcopy(model['spacing']))
    if not use_gpu:
        r = _cfg_get(cfg, 'cpu_model_spacing_increase_ratio', 1.0)
        model_spacing = [v * r for v in model_spacing]
    dev = next(model['net'].parameters()).device
    identity = is_identity_resample(image, model_spacing, model['max_stride'])
    if identity:
        iso_image = image
        data = image.data
        vol = (data if torch.is_tensor(data) else torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)))
        vol = vol.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    else:
        # resample to the model spacing on the device (:267)
        with torch.cuda.device(dev):
            iso_image = resample_spacing(Image3d(image.data if not torch.is_tensor(image.data) else image.data.to(dev),
                                                 image.GetSpacing(), image.GetOrigin(), image.GetDirection()),
                                         model_spacing, model['max_stride'], model['interpolation'])
        vol = iso_image.data
    if bbox_start_voxel is not None and bbox_end_voxel is not None:
        # convert the bounding box to the iso image frame (:292-302)
        if not identity:
            bs = image.TransformContinuousIndexToPhysicalPoint([float(v) for v in bbox_start_voxel])
            be = image.TransformContinuousIndexToPhysicalPoint([float(v) for v in bbox_end_voxel])
            bbox_start_voxel = iso_image.TransformPhysicalPointToIndex(bs)
            bbox_end_voxel = iso_image.TransformPhysicalPointToIndex(be)
        bbox_start_voxel = [max(0, int(v)) for v in bbox_start_voxel]
        bbox_end_voxel = [min(int(bbox_end_voxel[a]), iso_image.GetSize()[a]) for a in range(3)]
    acc, mask = segmentation_volume_device(model, cfg, vol, bbox_start_voxel=bbox_start_voxel,
                                           bbox_end_voxel=bbox_end_voxel, use_gpu=use_gpu, spacing=model_spacing)
    if not identity:
        # resample every class map back to the scan's grid (pad 1.0 for the background class, 0.0 otherwise) and take
        # the first-argmax there (:329-338)
        num_classes = model['out_channels']
        X, Y, Z = image.GetSize()
        back = torch.empty((num_classes, Z, Y, X), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            for c in range(num_classes):
                iso_c = Image3d(acc[c], iso_image.GetSpacing(), iso_image.GetOrigin(), iso_image.GetDirection())
                back[c] = resample(iso_c, image, 'LINEAR', 1.0 if c == 0 else 0.0).data
            ones = [np.ones(n, dtype=np.int32) for n in (X, Y, Z)]
            mask = model['engine'].finalize(back, ones)
        acc = back
    num_classes = model['out_channels']
    mean_probs = []
    for c in range(num_classes):
        im = Image3d(acc[c])
        im.CopyInformation(image)
        mean_probs.append(im)
    labels = list(range(1, num_classes))
    if _cfg_get(cfg, 'pick_largest_cc', False):                      # :341-342
        mask = _cc_filter_device(mask, labels, 0)
    if _cfg_get(cfg, 'remove_small_cc', 0) > 0:                      # :345-347
        mask = _cc_filter_device(mask, labels, int(cfg['remove_small_cc']))
    mask_im = Image3d(mask)
    mask_im.CopyInformation(image)
    return mean_probs, mask_im


def segmentation(input_path, model_folder, output_folder, seg_name, gpu_id, return_mask, save_mask, save_image, save_prob):
    """Volumetric segmentation engine: same inputs/outputs as the reference (:353-493).
    Launched as `python -m torch.distributed.run --nproc-per-node N -m segmentation3d.seg_infer ...`, every process takes
    the GPU of its local rank and the cases rank::N of the list (returned masks: this rank's cases, in list order)."""
    rank, world, local = launch_rank()
    requested_gpu_id = gpu_id
    if world > 1:
        gpu_id = local
    begin = time.time()
    models = load_models(model_folder, gpu_id)
    load_model_time = time.time() - begin

    is_dicom_folder = False
    if os.path.isfile(input_path):
        if input_path.endswith('.txt'):
            file_name_list, file_path_list = read_test_txt(input_path)
        elif input_path.endswith(('.mhd', '.mha', '.nii.gz', '.nii', '.hdr', '.image3d')):
            file_name_list, file_path_list = [os.path.basename(input_path)], [input_path]
        else:
            raise ValueError('Unsupported input path.')
    elif os.path.isdir(input_path):
        if glob.glob(os.path.join(input_path, '*.dcm')):
            raise NotImplementedError('DICOM series input needs SimpleITK/GDCM, which this build does not bundle')
        file_name_list, file_path_list = read_test_folder(input_path, is_dicom_folder)
        if len(file_name_list) == 0:
            raise ValueError('Empty test folder!')
    else:
        raise ValueError('The file {} does not exist.'.format(input_path))

    file_name_list, file_path_list = shard_case_list(file_name_list, file_path_list, rank, world)
    infer_cfg = models['infer_cfg']
    scale = infer_cfg.general.single_scale
    # reference quirk kept (core/seg_infer.py:420): only selects the cpu_*_ratio knobs; every rank of a sharded launch
    # follows the -g value the user passed, so all cases are segmented with the same partition settings
    use_gpu = requested_gpu_id > 0
    masks, total_inference_time, num_success_case = [], 0, 0
    # host I/O overlaps the GPU: the next case is read while this one is segmented, results are compressed and written in
    # the background (SEG3D_IO_THREADS=0: strictly serial, as the reference)
    writer = AsyncImageWriter(io_threads())
    images = prefetch_images(file_path_list, np.float32, enabled=io_threads() > 0)
    try:
        for i, file_path in enumerate(file_path_list):
            print('{}: {}'.format(i, file_path))
            image, read_image_time = next(images)

            begin = time.time()
            if scale == 'coarse':
                mean_probs, mask = segmentation_volume(models['coarse_model'], infer_cfg.coarse, image, None, None, use_gpu)
            elif scale == 'fine':
                mean_probs, mask = segmentation_volume(models['fine_model'], infer_cfg.fine, image, None, None, use_gpu)
            elif scale == 'DISABLE':
                print('Coarse segmentation: ')
                _, mask = segmentation_volume(models['coarse_model'], infer_cfg.coarse, image, None, None, use_gpu)
                start_voxel, end_voxel = get_bounding_box(mask, None)
                bbox_ratio = 100
                for a in range(3):
                    bbox_ratio *= (end_voxel[a] - start_voxel[a]) / mask.GetSize()[a]
                print('Fine segmentation (bbox ratio: {:.2f}%): '.format(bbox_ratio))
                mean_probs, mask = segmentation_volume(models['fine_model'], infer_cfg.fine, image, start_voxel, end_voxel, use_gpu)
            else:
                raise ValueError('Unsupported scale type!')
            torch.cuda.synchronize()
            inference_time = time.time() - begin
            if return_mask:
                masks.append(mask)

            begin = time.time()
            case_name = file_name_list[i]
            if save_mask or save_image or save_prob:
                os.makedirs(os.path.join(output_folder, case_name), exist_ok=True)
            if save_mask:
                writer.write(mask, os.path.join(output_folder, case_name, seg_name), True)
            if save_image:
                writer.write(image, os.path.join(output_folder, case_name, 'org.mha'), True)
            if save_prob:
                for c, prob in enumerate(mean_probs):
                    writer.write(prob, os.path.join(output_folder, case_name, 'mean_prob_{}.mha'.format(c)), True)
            save_time = time.time() - begin

            total_test_time = load_model_time + read_image_time + inference_time + save_time
            total_inference_time += inference_time
            num_success_case += 1
            print('total test time: {:.2f}, average inference time: {:.2f}'.format(
                total_test_time, total_inference_time / num_success_case))
    finally:
        images.close()
        writer.close()
    return masks
