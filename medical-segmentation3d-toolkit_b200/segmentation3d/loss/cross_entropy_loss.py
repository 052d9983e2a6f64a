"""Cross entropy on the network output (drop-in for reference loss/cross_entropy_loss.py:5-18, `loss.name = 'CE'` in
core/seg_train.py:98-99).  Same constructor as the reference's wrapper of nn.CrossEntropyLoss - (weight, size_average,
ignore_index, reduce, reduction), the two deprecated flags folding into `reduction` the way torch folds them - and the
same forward: the target's channel axis is squeezed and cast to class indices, the input (the network's probabilities,
which the reference feeds in as logits) goes through log-softmax + NLL.  Fused CUDA forward/backward (csrc/loss.cu)."""
import torch
from torch import nn

from segmentation3d.loss._kernels import CrossEntropyFunction


class CrossEntropyLoss(nn.Module):
    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction='mean'):
        super(CrossEntropyLoss, self).__init__()
        if size_average is not None or reduce is not None:          # torch.nn._reduction.legacy_get_string
            size_average = True if size_average is None else size_average
            reduce = True if reduce is None else reduce
            reduction = ('mean' if size_average else 'sum') if reduce else 'none'
        if reduction not in ('mean', 'sum', 'none'):
            raise ValueError('{} is not a valid value for reduction'.format(reduction))
        self.weight = None if weight is None else torch.as_tensor(weight, dtype=torch.float32)
        self.ignore_index, self.reduction = ignore_index, reduction

    def forward(self, input, target):
        assert isinstance(input, torch.Tensor)
        assert isinstance(target, torch.Tensor)
        if target.dim() == input.dim():
            target = torch.squeeze(target, dim=1)
        if input.dim() == 2:        # [sample, class] -> one "batch" with the samples as voxels
            input = input.t().unsqueeze(0)
            out = CrossEntropyFunction.apply(input, target, self.weight, self.ignore_index, self.reduction)
            return out.view(-1) if self.reduction == 'none' else out
        return CrossEntropyFunction.apply(input, target, self.weight, self.ignore_index, self.reduction)
