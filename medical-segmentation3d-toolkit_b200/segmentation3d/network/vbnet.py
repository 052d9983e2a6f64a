"""VBNet plug-in (drop-in for reference segmentation3d/network/vbnet.py:11-53): VNet topology
with bottleneck residual blocks (ratio 4) in down_64, down_128, down_256, up_256, up_128."""
from segmentation3d.network._graph import VShapedNet
from segmentation3d.network.module.weight_init import kaiming_weight_init, gaussian_weight_init


def parameters_kaiming_init(net):
    net.apply(kaiming_weight_init)


def parameters_gaussian_init(net):
    net.apply(gaussian_weight_init)


class SegmentationNet(VShapedNet):
    arch = 'vbnet'
    bottleneck_stages = ('down_64', 'down_128', 'down_256', 'up_256', 'up_128')
