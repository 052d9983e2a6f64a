"""Focal loss on probabilities (drop-in for reference loss/focal_loss.py:5-61); fused CUDA forward/backward.
Input [B,C,(D,)H,W] or [N,C]; target holds class indices (any float/int dtype), one per voxel."""
import torch
from torch import nn

from segmentation3d.loss._kernels import FocalFunction


class FocalLoss(nn.Module):
    def __init__(self, class_num, alpha=None, gamma=2, size_average=True, use_gpu=True):
        super(FocalLoss, self).__init__()
        if alpha is None:
            self.alpha = torch.ones(class_num, 1) / class_num
        else:
            assert len(alpha) == class_num
            a = torch.FloatTensor(alpha).unsqueeze(1)
            self.alpha = a / a.sum()
        if use_gpu:
            self.alpha = self.alpha.cuda()
        self.gamma, self.class_num, self.size_average = gamma, class_num, size_average

    def forward(self, input, target):
        assert input.dim() in (2, 4, 5)
        if input.dim() == 2:        # [sample, class] -> one "batch" with samples as voxels
            input = input.t().unsqueeze(0)
        return FocalFunction.apply(input, target, self.alpha, self.gamma, self.size_average)
