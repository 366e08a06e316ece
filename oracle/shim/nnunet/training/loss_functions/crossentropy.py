"""ORACLE (test infrastructure). Restates nnunet@77bc485 training/loss_functions/crossentropy.py
(SURVEY.md Appendix A): CrossEntropyLoss that accepts a (B,1,...) float target."""
from torch import nn


class RobustCrossEntropyLoss(nn.CrossEntropyLoss):
    def forward(self, input, target):
        if len(target.shape) == len(input.shape):
            assert target.shape[1] == 1
            target = target[:, 0]
        return super().forward(input, target.long())
