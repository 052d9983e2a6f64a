"""Drop-in for reference network/module/init.py:4-21 (the older home of kaiming_weight_init)."""
from segmentation3d.network.module.weight_init import kaiming_weight_init  # noqa: F401
