"""ORACLE (test infrastructure). Restates nnunet@77bc485 nnunet/utilities/tensor_utilities.py
``sum_tensor`` (sum over sorted axes, descending when not keepdim)."""
import numpy as np


def sum_tensor(inp, axes, keepdim=False):
    axes = np.unique(axes).astype(int)
    if keepdim:
        for ax in axes:
            inp = inp.sum(int(ax), keepdim=True)
    else:
        for ax in sorted(axes, reverse=True):
            inp = inp.sum(int(ax))
    return inp
