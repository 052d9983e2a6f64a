"""VNet plug-in (drop-in for reference segmentation3d/network/vnet.py:10-51).

Discovered by `importlib.import_module('segmentation3d.network.' + name)`; exposes
SegmentationNet(in_channels, out_channels), parameters_kaiming_init, parameters_gaussian_init.
"""
from segmentation3d.network._graph import VShapedNet
from segmentation3d.network.module.weight_init import kaiming_weight_init, gaussian_weight_init


def parameters_kaiming_init(net):
    net.apply(kaiming_weight_init)


def parameters_gaussian_init(net):
    net.apply(gaussian_weight_init)


class SegmentationNet(VShapedNet):
    """16-32-64-128-256 encoder/decoder with plain residual blocks; probabilities out."""
    arch = 'vnet'
    bottleneck_stages = ()
