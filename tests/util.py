"""Shared helpers of the parity tests (oracle side + CUDA side)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lifelong-nnunet_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def rel_err(a, b):
    """||a-b||_inf / max(||b||_inf, 1e-6)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def oracle_net(geom, seed=0):
    from oracle import step
    return step.build_network(geom.in_channels, geom.base_features, geom.num_classes, [list(k) for k in geom.pool], seed=seed,
                              max_num_features=geom.max_features)


def cuda_net(geom, state_dict=None, precision="fp32"):
    from b200unet.generic_UNet import Generic_UNet
    net = Generic_UNet(geom.in_channels, geom.base_features, geom.num_classes, geom.num_pool,
                       pool_op_kernel_sizes=[list(k) for k in geom.pool],
                       conv_kernel_sizes=[[3, 3, 3]] * (geom.num_pool + 1), max_num_features=geom.max_features)
    if state_dict is not None:
        net.load_state_dict(state_dict)
    net.precision = precision
    return net.cuda()
