"""ORACLE (test infrastructure). Restates nnunet@77bc485 training/loss_functions/dice_loss.py
(SURVEY.md Appendix A): get_tp_fp_fn_tn, SoftDiceLoss, DC_and_CE_loss.  The reference restates the hard
tp/fp/fn variant at nnUNetTrainerMultiHead.py:938-951 and builds the loss at :1385
(``DC_and_CE_loss({'batch_dice':..., 'smooth':1e-5, 'do_bg':False}, {})``)."""
import numpy as np
import torch
from torch import nn

from nnunet.training.loss_functions.crossentropy import RobustCrossEntropyLoss
from nnunet.utilities.nd_softmax import softmax_helper
from nnunet.utilities.tensor_utilities import sum_tensor


def get_tp_fp_fn_tn(net_output, gt, axes=None, mask=None, square=False):
    if axes is None:
        axes = tuple(range(2, len(net_output.size())))
    shp_x, shp_y = net_output.shape, gt.shape
    with torch.no_grad():
        if len(shp_x) != len(shp_y):
            gt = gt.view((shp_y[0], 1, *shp_y[1:]))
        if all(i == j for i, j in zip(net_output.shape, gt.shape)):
            y_onehot = gt
        else:
            gt = gt.long()
            y_onehot = torch.zeros(shp_x, device=net_output.device)
            y_onehot.scatter_(1, gt, 1)
    tp = net_output * y_onehot
    fp = net_output * (1 - y_onehot)
    fn = (1 - net_output) * y_onehot
    tn = (1 - net_output) * (1 - y_onehot)
    if mask is not None:
        tp, fp, fn, tn = [torch.stack(tuple(x_i * mask[:, 0] for x_i in torch.unbind(t, dim=1)), dim=1)
                          for t in (tp, fp, fn, tn)]
    if square:
        tp, fp, fn, tn = tp ** 2, fp ** 2, fn ** 2, tn ** 2
    if len(axes) > 0:
        tp, fp, fn, tn = [sum_tensor(t, axes, keepdim=False) for t in (tp, fp, fn, tn)]
    return tp, fp, fn, tn


class SoftDiceLoss(nn.Module):
    def __init__(self, apply_nonlin=None, batch_dice=False, do_bg=True, smooth=1.):
        super().__init__()
        self.do_bg, self.batch_dice, self.apply_nonlin, self.smooth = do_bg, batch_dice, apply_nonlin, smooth

    def forward(self, x, y, loss_mask=None):
        shp_x = x.shape
        axes = [0] + list(range(2, len(shp_x))) if self.batch_dice else list(range(2, len(shp_x)))
        if self.apply_nonlin is not None:
            x = self.apply_nonlin(x)
        tp, fp, fn, _ = get_tp_fp_fn_tn(x, y, axes, loss_mask, False)
        nominator = 2 * tp + self.smooth
        denominator = 2 * tp + fp + fn + self.smooth
        dc = nominator / (denominator + 1e-8)
        if not self.do_bg:
            dc = dc[1:] if self.batch_dice else dc[:, 1:]
        return -dc.mean()


class DC_and_CE_loss(nn.Module):
    def __init__(self, soft_dice_kwargs, ce_kwargs, aggregate="sum", square_dice=False, weight_ce=1, weight_dice=1,
                 log_dice=False, ignore_label=None):
        super().__init__()
        assert not square_dice and not log_dice and ignore_label is None, "not on the hot path"
        self.weight_dice, self.weight_ce, self.aggregate = weight_dice, weight_ce, aggregate
        self.ce = RobustCrossEntropyLoss(**ce_kwargs)
        self.dc = SoftDiceLoss(apply_nonlin=softmax_helper, **soft_dice_kwargs)

    def forward(self, net_output, target):
        dc_loss = self.dc(net_output, target) if self.weight_dice != 0 else 0
        ce_loss = self.ce(net_output, target[:, 0].long()) if self.weight_ce != 0 else 0
        if self.aggregate == "sum":
            return self.weight_ce * ce_loss + self.weight_dice * dc_loss
        raise NotImplementedError("nah son")
