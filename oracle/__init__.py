"""CPU oracle for the volumetric-segmentation hot path (TEST INFRASTRUCTURE ONLY).

This package is a plain CPU restatement (torch fp32 functional ops + numpy) of the
reference algorithm on the path named by BASELINE.json `north_star`:

  * VNet / VBNet forward          -> oracle/net.py            (reference: segmentation3d/network/*.py)
  * weight initialisation         -> oracle/init.py           (reference: network/module/weight_init.py)
  * Dice / focal losses           -> oracle/loss.py           (reference: segmentation3d/loss/*.py)
  * patch grid, normalisers,
    sliding-window blend, argmax  -> oracle/sliding_window.py (reference: core/seg_infer.py, utils/image_tools.py)
  * Dice-ratio metric             -> oracle/metrics.py        (reference: utils/metrics.py)
  * the same network program with the product's half-precision rounding points injected (not a restatement of the
    reference, which is fp32: the checker for the bf16 training path)   -> oracle/reduced_precision.py

The arithmetic of the reference lives in a third-party dependency, **torch**
(reference README pins Pytorch=1.3.0; this image has torch 2.11.0+cu128): conv3d /
conv_transpose3d / group_norm / softmax.  The oracle calls exactly those ATen CPU
fp32 ops through `torch.nn.functional`, in the order the reference modules call them.

PARITY PINNING: the reference's own tests assert no numeric value, so the
oracle is pinned against outputs of the reference itself, run in the build container by
`tests/golden/make_golden.py` (imports /root/reference read-only under small stand-ins for
the absent SimpleITK / easydict) and committed under `tests/golden/`.  `tests/test_oracle_*.py`
check the oracle against every fixture.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package, and only as the checker / the timed CPU baseline.  Nothing in
`medical-segmentation3d-toolkit_b200/` imports it; the product fails loudly without its
CUDA library.
"""
