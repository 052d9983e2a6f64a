"""Initialisers applied with `net.apply(...)` (reference network/module/weight_init.py:4-29).
They dispatch on the class NAME, exactly like the reference: anything whose name contains
'Conv3d' / 'ConvTranspose3d' gets its weight re-drawn and its bias zeroed; GroupNorm holders are
left at gamma=1, beta=0."""
import torch.nn as nn


def _is_conv(m):
    name = type(m).__name__
    return 'Conv3d' in name or 'ConvTranspose3d' in name


def kaiming_weight_init(m, bn_std=0.02):
    name = type(m).__name__
    if _is_conv(m) or 'Linear' in name:
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            m.bias.data.zero_()
    elif 'BatchNorm' in name:
        m.weight.data.normal_(1.0, bn_std)
        m.bias.data.zero_()


def gaussian_weight_init(m, conv_std=0.01, bn_std=0.01):
    name = type(m).__name__
    if _is_conv(m):
        m.weight.data.normal_(0, conv_std)
        if m.bias is not None:
            m.bias.data.zero_()
    elif 'BatchNorm' in name:
        m.weight.data.normal_(1.0, bn_std)
        m.bias.data.zero_()
