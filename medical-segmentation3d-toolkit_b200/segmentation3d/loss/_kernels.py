"""autograd bridges from the loss modules to the fused CUDA reductions (csrc/loss.cu)."""
import torch

from segmentation3d._b200 import lib


def _check(probs, target):
    if not probs.is_cuda:
        raise RuntimeError('segmentation3d (B200 build) losses run on CUDA tensors only (no CPU fallback)')
    B, C = probs.shape[0], probs.shape[1]
    n = probs[0, 0].numel()
    if target.numel() != B * n:
        raise ValueError('target must hold one label per voxel: got %s for input %s' % (tuple(target.shape), tuple(probs.shape)))
    return B, C, n


class DiceFunction(torch.autograd.Function):
    """sum_c w_c * mean_b [ 1 - (2 I + eps) / (A + T + eps) ],  I = sum q t, A = sum q^2, T = sum t^2,
    q = p_c [p_c > 1/C], t = [target == c]  (closed form of loss/multi_dice_loss.py:24-43 +
    loss/binary_dice_loss.py:9-36).  One kernel pass forward, one backward."""
    EPS = 1e-6

    @staticmethod
    def forward(ctx, probs, target, weights):
        B, C, n = _check(probs, target)
        p = probs.detach().contiguous().float()
        t = target.detach().contiguous().float()
        terms = torch.zeros((B, C, 3), dtype=torch.float64, device=p.device)
        lib.call('seg3d_dice_terms', lib.ptr(p), lib.ptr(t), B, C, n, lib.ptr(terms), lib.stream_ptr())
        I, A, T = terms[..., 0], terms[..., 1], terms[..., 2]
        S = A + T + DiceFunction.EPS
        per = 1.0 - (2.0 * I + DiceFunction.EPS) / S                   # [B, C]
        w = weights.to(device=p.device, dtype=torch.float64)
        loss = (per.mean(0) * w).sum()
        ctx.save_for_backward(p, t, I, S, w)
        return loss.float()

    @staticmethod
    def backward(ctx, gout):
        p, t, I, S, w = ctx.saved_tensors
        B, C = p.shape[0], p.shape[1]
        n = p[0, 0].numel()
        g = gout.double()
        coef = torch.empty((B, C, 2), dtype=torch.float64, device=p.device)
        coef[..., 0] = -2.0 * w.view(1, C) / (B * S) * g
        coef[..., 1] = 2.0 * w.view(1, C) * (2.0 * I + DiceFunction.EPS) / (B * S * S) * g
        coef = coef.float().contiguous()
        grad = torch.empty_like(p)
        lib.call('seg3d_dice_bwd', lib.ptr(p), lib.ptr(t), B, C, n, lib.ptr(coef), lib.ptr(grad), lib.stream_ptr())
        return grad, None, None


class FocalFunction(torch.autograd.Function):
    """-alpha_t (1 - p_t)^gamma log(p_t + 1e-10), mean or sum over voxels (loss/focal_loss.py:27-61)."""

    @staticmethod
    def forward(ctx, probs, target, alpha, gamma, size_average):
        B, C, n = _check(probs, target)
        p = probs.detach().contiguous().float()
        t = target.detach().contiguous().float()
        a = alpha.to(device=p.device, dtype=torch.float32).contiguous().view(-1)
        if a.numel() != C:
            raise ValueError('alpha must hold one value per class')
        part = torch.zeros((2,), dtype=torch.float64, device=p.device)
        lib.call('seg3d_focal_fwd', lib.ptr(p), lib.ptr(t), B, C, n, lib.ptr(a), float(gamma), lib.ptr(part), lib.stream_ptr())
        bad = int(part[1].item())
        if bad:          # the reference's one-hot gather (loss/focal_loss.py:46-48) raises on such a target
            raise IndexError('FocalLoss: %d target voxels carry a label outside [0, %d)' % (bad, C))
        ctx.save_for_backward(p, t, a)
        ctx.gamma, ctx.scale = float(gamma), (1.0 / (B * n) if size_average else 1.0)
        return (part[0] * ctx.scale).float()

    @staticmethod
    def backward(ctx, gout):
        p, t, a = ctx.saved_tensors
        B, C = p.shape[0], p.shape[1]
        n = p[0, 0].numel()
        grad = torch.empty_like(p)
        lib.call('seg3d_focal_bwd', lib.ptr(p), lib.ptr(t), B, C, n, lib.ptr(a), ctx.gamma, float(gout) * ctx.scale,
                 lib.ptr(grad), lib.stream_ptr())
        return grad, None, None, None, None


class CrossEntropyFunction(torch.autograd.Function):
    """nn.CrossEntropyLoss on [B,C,*spatial] inputs with one class index per voxel (loss/cross_entropy_loss.py:5-18):
    w_t (logsumexp_c x_c - x_t), reduced by 'mean' (sum / sum of w_t), 'sum' or 'none'; ignore_index voxels drop out."""

    @staticmethod
    def forward(ctx, x, target, weight, ignore_index, reduction):
        B, C, n = _check(x, target)
        xc = x.detach().contiguous().float()
        t = target.detach().contiguous().float()
        w = None if weight is None else weight.to(device=xc.device, dtype=torch.float32).contiguous().view(-1)
        if w is not None and w.numel() != C:
            raise ValueError('weight must hold one value per class')
        part = torch.zeros((2,), dtype=torch.float64, device=xc.device)
        lmap = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.float32, device=xc.device) if reduction == 'none' else None
        lib.call('seg3d_ce_fwd', lib.ptr(xc), lib.ptr(t), B, C, n, lib.ptr(w), int(ignore_index), lib.ptr(part), lib.ptr(lmap),
                 lib.stream_ptr())
        ctx.save_for_backward(xc, t, part)
        ctx.w, ctx.ignore_index, ctx.reduction = w, int(ignore_index), reduction
        if reduction == 'none':
            return lmap
        if reduction == 'sum':
            return part[0].float()
        return (part[0] / part[1]).float()

    @staticmethod
    def backward(ctx, gout):
        xc, t, part = ctx.saved_tensors
        B, C = xc.shape[0], xc.shape[1]
        n = xc[0, 0].numel()
        grad = torch.empty_like(xc)
        gmap, scale = None, 1.0
        if ctx.reduction == 'none':
            gmap = gout.detach().contiguous().float()
        elif ctx.reduction == 'sum':
            scale = float(gout)
        else:
            scale = float(gout) / float(part[1])
        lib.call('seg3d_ce_bwd', lib.ptr(xc), lib.ptr(t), B, C, n, lib.ptr(ctx.w), ctx.ignore_index, scale, lib.ptr(gmap),
                 lib.ptr(grad), lib.stream_ptr())
        return grad, None, None, None, None
