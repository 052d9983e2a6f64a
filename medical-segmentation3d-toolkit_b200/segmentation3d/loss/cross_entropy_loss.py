"""Cross-entropy wrapper (drop-in for reference loss/cross_entropy_loss.py:5-18).  Like the reference it
hands the network's probabilities to torch.nn.CrossEntropyLoss; it is not on the north-star path and
delegates to torch."""
import torch.nn as nn


class CrossEntropyLoss(nn.Module):
    def __init__(self, weight=None, ignore_index=-100, reduction='mean'):
        super(CrossEntropyLoss, self).__init__()
        self.loss = nn.CrossEntropyLoss(weight=weight, ignore_index=ignore_index, reduction=reduction)

    def forward(self, input, target):
        if target.dim() == input.dim():
            target = target.squeeze(1)
        return self.loss(input, target.long())
