"""ORACLE (test infrastructure). Minimal stand-in for nnunet@77bc485 ``SegmentationNetwork`` /
``NeuralNetwork`` (only what the hot path touches: it is an nn.Module with a few attributes; the
sliding-window inference lives outside the hot path)."""
from torch import nn


class NeuralNetwork(nn.Module):
    def get_device(self):
        p = next(self.parameters())
        return "cpu" if p.device.type == "cpu" else p.device.index


class SegmentationNetwork(NeuralNetwork):
    def __init__(self):
        super().__init__()
        self.input_shape_must_be_divisible_by = None
        self.conv_op = None
        self.num_classes = None
        self.inference_apply_nonlin = lambda x: x
