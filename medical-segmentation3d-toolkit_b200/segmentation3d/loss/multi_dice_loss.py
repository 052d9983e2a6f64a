"""Multi-class Dice loss (drop-in for reference loss/multi_dice_loss.py:6-43): per class a binary Dice of
[p_i > 1/C] * p_i against [target == i], class-weighted with w / sum(w); fused CUDA reduction."""
import torch
import torch.nn as nn

from segmentation3d.loss._kernels import DiceFunction


class MultiDiceLoss(nn.Module):
    def __init__(self, weights, num_class, use_gpu):
        super(MultiDiceLoss, self).__init__()
        self.num_class = num_class
        assert len(weights) == self.num_class, "the length of weight must equal to num_class"
        w = torch.FloatTensor(weights)
        self.weights = w / w.sum()
        if use_gpu:
            self.weights = self.weights.cuda()

    def forward(self, input_tensor, target):
        assert input_tensor.size(1) == self.num_class
        return DiceFunction.apply(input_tensor, target, self.weights)
