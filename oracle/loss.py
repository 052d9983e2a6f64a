"""Oracle: Dice / focal losses on probabilities (torch CPU fp32, autograd-capable).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference (relative to /root/reference):
  segmentation3d/loss/binary_dice_loss.py:9-36
  segmentation3d/loss/multi_dice_loss.py:9-43
  segmentation3d/loss/focal_loss.py:7-61
  segmentation3d/loss/cross_entropy_loss.py:5-18
"""
import torch


def binary_dice_loss(inp, target):
    """binary_dice_loss.py:9-36.  inp [B,2,...], target [B,1,...] (float)."""
    b = inp.size(0)
    pred, label = inp.max(1)                       # :13  (ties -> index 0)
    pred = pred * label.float()                    # :14
    pred = pred.float().view(b, -1)
    tgt = target.float().view(b, -1)
    inter = torch.sum(pred * tgt, 1)               # :25
    area_p = torch.sum(pred * pred, 1)             # :26
    area_t = torch.sum(tgt * tgt, 1)               # :27
    eps = torch.tensor(1e-6)
    batch_loss = torch.tensor(1.0) - (torch.tensor(2.0) * inter + eps) / (area_p + area_t + eps)  # :33
    return batch_loss.mean()


def multi_dice_loss(probs, target, weights):
    """multi_dice_loss.py:24-43: per class, BinaryDice(cat[const 1/C, p_i], [t == i]), weighted by
    w / sum(w) (:19)."""
    num_class = probs.size(1)
    w = torch.tensor(weights, dtype=torch.float32)
    w = w / w.sum()
    total = 0
    for i in range(num_class):
        p_i = probs[:, i:i + 1]
        slice_i = torch.cat([1.0 / num_class + torch.zeros_like(p_i), p_i], dim=1)   # :36
        target_i = (target == i).float()                                             # :37
        total = total + binary_dice_loss(slice_i, target_i) * w[i]                   # :41
    return total


def multi_dice_terms(probs, target):
    """Closed form of the per-sample, per-class sums the loss is built from:
    q = p_i * [p_i > 1/C]; I = sum q t_i; A = sum q^2; T = sum t_i^2.  Returns [B,C,3] float64."""
    b, c = probs.shape[:2]
    out = torch.zeros(b, c, 3, dtype=torch.float64)
    for i in range(c):
        p = probs[:, i].reshape(b, -1).double()
        t = (target.reshape(b, -1) == i).double()
        q = p * (probs[:, i].reshape(b, -1) > (1.0 / c)).double()
        out[:, i, 0] = (q * t).sum(1)
        out[:, i, 1] = (q * q).sum(1)
        out[:, i, 2] = (t * t).sum(1)
    return out


def focal_loss(probs, target, class_num, alpha=None, gamma=2, size_average=True):
    """focal_loss.py:27-61 (input = probabilities, channels moved last, p_t + 1e-10)."""
    if alpha is None:
        a = torch.ones(class_num, 1) / class_num                    # :11
    else:
        a = torch.tensor(alpha, dtype=torch.float32).unsqueeze(1)
        a = a / a.sum()                                             # :14-16
    if probs.dim() == 4:
        x = probs.permute(0, 2, 3, 1).contiguous()
    elif probs.dim() == 5:
        x = probs.permute(0, 2, 3, 4, 1).contiguous()
    else:
        x = probs
    x = x.view(x.numel() // class_num, class_num)
    t = target.long().view(-1)
    mask = torch.eye(class_num)[t]
    al = a[t]
    p = (x * mask).sum(1).view(-1, 1) + 1e-10                       # :48
    logp = p.log()
    if gamma > 0:
        bl = -al * torch.pow(1 - p, gamma) * logp                   # :52
    else:
        bl = -al * logp
    return bl.mean() if size_average else bl.sum()


def cross_entropy_loss(inp, target, weight=None, ignore_index=-100, reduction='mean'):
    """cross_entropy_loss.py:12-18: squeeze the target's channel axis, cast to long, nn.CrossEntropyLoss on the input as
    given (the training loop feeds the network's probabilities in as logits, core/seg_train.py:98-99,122-123)."""
    import torch.nn.functional as F
    tgt = torch.squeeze(target, dim=1).long()
    return F.cross_entropy(inp, tgt, weight=weight, ignore_index=ignore_index, reduction=reduction)
