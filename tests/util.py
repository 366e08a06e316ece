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


# -- committed fixtures of the rows around the step (oracle/gen_golden_f.py) -------------------------------------------------------
def load_augment_fixture():
    """(cases, plan, patch, strides, gen_patch, params, data, targets, margin) of tests/golden/augment_tiny.npz"""
    import json
    import numpy as np
    from oracle import gen_golden_f
    z = np.load(os.path.join(ROOT, "tests", "golden", "augment_tiny.npz"))
    plan = json.loads(bytes(z["plan_json"]).decode())
    targets = [z["target%d" % k].astype(np.float32) for k in range(len(gen_golden_f.STRIDES))]
    return (gen_golden_f.augment_cases(), plan, gen_golden_f.PATCH, gen_golden_f.STRIDES, tuple(int(v) for v in z["gen_patch"]),
            gen_golden_f.AUG_PARAMS, z["data"], targets, z["margin"])


def augment_mismatch(data, targets, ref_data, ref_targets, margin):
    """(fraction of data voxels, clear of the crop border, that are off by more than 2e-3 of the value range; worst target agreement)"""
    import numpy as np
    scale = float(np.abs(ref_data).max())
    safe = np.broadcast_to((np.abs(margin) > 0.05)[:, None], ref_data.shape)
    bad = float((np.abs(np.asarray(data) - ref_data)[safe] > 2e-3 * scale).mean())
    agree = min(float((np.asarray(t) == r).mean()) for t, r in zip(targets, ref_targets))
    return bad, agree
