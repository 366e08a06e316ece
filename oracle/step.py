"""ORACLE (test infrastructure, CPU/PyTorch fp32) -- the reference's per-step training path restated on top of the
``nnunet`` shim (oracle/shim) so it can travel to the GPU box.

* network ctor call      -> reference nnunet_ext/training/network_training/nnViTUNetTrainer.py:101-138
                            (verbatim copy of nnUNetTrainerV2.initialize_network)
* optimizer              -> reference .../multihead/nnUNetTrainerMultiHead.py:294-301
* one iteration          -> reference .../multihead/nnUNetTrainerMultiHead.py:606-656 (fp32 branch :632-641)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import os
import sys

import torch
from torch import nn

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from nnunet.network_architecture.generic_UNet import Generic_UNet  # noqa: E402
from nnunet.network_architecture.initialization import InitWeights_He  # noqa: E402

from . import cl_losses  # noqa: E402


def build_network(input_channels, base_num_features, num_classes, pool_op_kernel_sizes, conv_kernel_sizes=None,
                  conv_per_stage=2, seed=0, max_num_features=None):
    """nnViTUNetTrainer.py:101-125 ctor call (InstanceNorm3d eps 1e-5 affine, Dropout p=0, LeakyReLU 1e-2,
    deep supervision, identity final nonlin, He init, conv pooling + conv upsampling)."""
    if conv_kernel_sizes is None:
        conv_kernel_sizes = [[3, 3, 3]] * (len(pool_op_kernel_sizes) + 1)
    torch.manual_seed(seed)
    net = Generic_UNet(input_channels, base_num_features, num_classes, len(pool_op_kernel_sizes), conv_per_stage, 2,
                       nn.Conv3d, nn.InstanceNorm3d, {'eps': 1e-5, 'affine': True}, nn.Dropout3d,
                       {'p': 0, 'inplace': True}, nn.LeakyReLU, {'negative_slope': 1e-2, 'inplace': True},
                       True, False, lambda x: x, InitWeights_He(1e-2), pool_op_kernel_sizes, conv_kernel_sizes,
                       False, True, True, max_num_features)
    return net


def make_optimizer(net, lr=1e-2, weight_decay=3e-5):
    """MultiHead:294-301: SGD(lr 1e-2, wd 3e-5, momentum 0.99, nesterov)."""
    return torch.optim.SGD(net.parameters(), lr, weight_decay=weight_decay, momentum=0.99, nesterov=True)


def run_iteration(net, optimizer, data, target, loss_fn, do_backprop=True, clip=12.0):
    """MultiHead:606-656, fp32 branch.  ``loss_fn(output, target)`` returns a 0-dim tensor.  Returns
    (loss value, output tuple)."""
    optimizer.zero_grad()
    output = net(data)
    l = loss_fn(output, target)
    if do_backprop:
        l.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        optimizer.step()
    return float(l.detach()), output


def base_loss_fn(weights, batch_dice=False):
    return lambda out, tgt: cl_losses.multiple_output_loss2(
        out, tgt, weights, loss=lambda x, y: cl_losses.dc_and_ce(x, y, batch_dice=batch_dice))


def ewc_loss_fn(net, weights, fisher, params, ewc_lambda, strict_reference=True, batch_dice=False):
    base = base_loss_fn(weights, batch_dice)
    return lambda out, tgt: base(out, tgt) + cl_losses.ewc_penalty(
        list(net.named_parameters()), fisher, params, ewc_lambda, strict_reference)


def rw_loss_fn(net, weights, fisher, params, scores, rw_lambda, strict_reference=True, batch_dice=False):
    base = base_loss_fn(weights, batch_dice)
    return lambda out, tgt: base(out, tgt) + cl_losses.rw_penalty(
        list(net.named_parameters()), fisher, params, scores, rw_lambda, strict_reference)
