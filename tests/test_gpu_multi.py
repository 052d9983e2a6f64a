"""Multi-GPU parity under `pytest -m gpu` (skipped on a one-GPU box): two ranks over NCCL, real kernels.

  tools/check_patch_shard.py        ONE volume, patches dealt over the ranks: all three exchange steps (all-reduce of the
                                    probability maps, reduce-scatter of slabs + mask all-gather, label max all-reduce) against
                                    the single-rank pass (reference loop core/seg_infer.py:313-327)
  tools/check_overlap_allreduce.py  data-parallel training (reference core/seg_train.py:77): gradient all-reduce overlapped with
                                    the backward pass vs issued after it, and the 2-rank step vs one process on the global batch
Their output is kept under gpurun_out/ (copied to profiles/ by hand)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _two_ranks(script, port):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tools', script)]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    out = os.path.join(ROOT, 'gpurun_out')
    if os.path.isdir(out):
        with open(os.path.join(out, 'multi_%s.log' % script[:-3]), 'w') as f:
            f.write(r.stdout)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def test_patch_sharded_volume_matches_single_rank():
    out = _two_ranks('check_patch_shard.py', 29531)
    assert out.count('label exchange mask agreement') == 2


def test_data_parallel_training_matches_global_batch():
    out = _two_ranks('check_overlap_allreduce.py', 29532)
    assert out.count('data-parallel (world 2') == 2
