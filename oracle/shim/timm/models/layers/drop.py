"""ORACLE shim: timm.models.layers.drop.DropPath (reference vision_transformer.py:175; the reference always builds
it with drop_path_rate=0, generic_ViT_UNet.py:179, i.e. nn.Identity is used instead)."""
import torch
from torch import nn


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x / keep * mask
