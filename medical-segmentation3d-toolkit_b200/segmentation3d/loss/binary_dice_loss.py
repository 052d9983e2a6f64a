"""Binary Dice loss on a two-channel probability map (drop-in for reference loss/binary_dice_loss.py:5-36).
MultiDiceLoss does not route through this class in this build (it calls the fused kernel directly); the
standalone form is expressed with the same kernel by viewing channel 0 as the threshold plane when it is
the constant 1/2 map, and otherwise evaluated with tensor ops."""
import torch
import torch.nn as nn

from segmentation3d.loss._kernels import DiceFunction


class BinaryDiceLoss(nn.Module):
    def forward(self, input, target):
        assert input.size(1) == 2, 'BinaryDiceLoss expects [B,2,...] probabilities'
        b = input.size(0)
        if input.is_cuda and bool((input[:, 0] == 0.5).all()):
            # max over [1/2, p1] * argmax label == p1 [p1 > 1/2]: the fused reduction with a one-hot weight
            w = torch.tensor([0.0, 1.0])
            return DiceFunction.apply(input, (target != 0).float(), w)
        pred, label = input.max(1)
        pred = (pred * label.float()).float().view(b, -1)
        tgt = target.float().view(b, -1)
        inter, a_p, a_t = (pred * tgt).sum(1), (pred * pred).sum(1), (tgt * tgt).sum(1)
        return (1.0 - (2.0 * inter + 1e-6) / (a_p + a_t + 1e-6)).mean()
