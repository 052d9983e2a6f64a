"""Dry run of GPU tests on the CPU: `.cuda()` is the identity and every C-ABI call goes to the emulations of the wiring tests
(tests/test_plan_wiring.py, test_sliding_wiring.py, test_device_crops_wiring.py).  It validates TEST CODE and host code - shapes,
API use, tolerances in exact arithmetic - before GPU time is spent on them; it says nothing about the kernels.

    python tests/dryrun_gpu_tests_on_cpu.py [tests/test_gpu_blocks.py ...]      (default: the gated tests of test_gpu_blocks.py)

Round 1: all 8 gated tests of tests/test_gpu_blocks.py pass this dry run (they still have to see a GPU)."""
import contextlib
import os
import sys

os.environ['SEG3D_TEST_UNVERIFIED'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, ROOT + '/medical-segmentation3d-toolkit_b200', ROOT + '/tests'):
    sys.path.insert(0, p)
import pytest
import torch

class _MP(object):
    def setattr(self, obj, name, value): setattr(obj, name, value)

class Plugin(object):
    @pytest.fixture(autouse=True)
    def _emulate(self, monkeypatch):
        import test_autograd_wiring as taw
        import test_device_crops_wiring as tdc
        import test_sliding_wiring as tsw
        from segmentation3d._b200 import lib, blocks
        from segmentation3d.utils import image3d
        tables = []
        real = {}
        def grab(install, *a):
            install(monkeypatch, *a)
            tables.append(lib.call)
        grab(taw._install, [])          # forward + backward entry points of the network, loss reductions
        grab(tsw._install, [])
        grab(tdc._install)
        def call(name, *a):
            last = None
            for t in tables:
                try:
                    return t(name, *a)
                except KeyError as e:
                    last = e
            raise last
        monkeypatch.setattr(lib, 'call', call)
        class P(object):
            def __init__(self, t, off): self.t, self.off = t, off
            def flat(self): return self.t.reshape(-1)[self.off:]
        monkeypatch.setattr(lib, 'ptr', lambda t, off=0: None if t is None else P(t, off))
        monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
        monkeypatch.setattr(torch.nn.Module, 'cuda', lambda self, *a, **k: self)
        monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
        monkeypatch.setattr(torch.cuda, 'device', lambda d: contextlib.nullcontext())
        monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
        monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
        monkeypatch.setattr(torch.cuda, 'device_count', lambda: 1)
        for name in ('set_device', 'manual_seed', 'manual_seed_all'):
            monkeypatch.setattr(torch.cuda, name, lambda *a, **k: None)
        monkeypatch.setattr(torch.cuda, 'current_device', lambda: 0)
        monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: False)
        monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self, *a, **k: self)
        monkeypatch.setattr(torch.Tensor, 'is_pinned', lambda self, *a, **k: False)
        from segmentation3d.core import seg_infer
        from segmentation3d.dataloader import device_loader
        real_init = device_loader.DeviceCropLoader.__init__
        monkeypatch.setattr(device_loader.DeviceCropLoader, '__init__',
                            lambda self, dataset, sampler, batch_size, device=None, cache_gb=None: real_init(self, dataset, sampler, batch_size, 'cpu', cache_gb))
        monkeypatch.setattr(seg_infer, '_device_for', lambda gpu_id: torch.device('cpu'))
        yield

targets = sys.argv[1:] or [os.path.join(ROOT, 'tests', 'test_gpu_blocks.py')]
sys.exit(pytest.main(['-q', '-m', 'gpu', '-p', 'no:cacheprovider'] + targets, plugins=[Plugin()]))
