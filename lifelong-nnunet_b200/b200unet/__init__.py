"""b200unet -- B200-native (sm_100a) implementation of Lifelong-nnUNet's per-step training hot path.

Host side of the C ABI in include/b2unet.h: the ``Generic_UNet`` plugin surface, the ``MultipleOutputLoss*`` classes
and the hot-path methods of the ``nnUNetTrainer{EWC,RW,LWF,MiB,PLOP,POD}`` trainers.
"""
from . import _lib  # noqa: F401
from .configs import CONFIGS, UNetGeometry  # noqa: F401
