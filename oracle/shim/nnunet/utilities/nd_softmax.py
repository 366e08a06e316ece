"""ORACLE (test infrastructure). Restates nnunet@77bc485 nnunet/utilities/nd_softmax.py
(``softmax_helper = lambda x: F.softmax(x, 1)``)."""
import torch.nn.functional as F


def softmax_helper(x):
    return F.softmax(x, 1)
