"""Mirror of MCAcc/interp2x_boundary3d.py:9-30 (autograd wrapper + module around the 2x upsampler with boundary flags)."""
import torch.nn as nn
from torch.autograd import Function

from .. import ops


class Interp2xBoundary3dFunction(Function):
    @staticmethod
    def forward(ctx, input, balance_value):
        output, is_boundary = ops.interp2x_boundary3d_forward(input.contiguous(), balance_value, 0)
        ctx.mark_non_differentiable(is_boundary)
        return output, is_boundary

    @staticmethod
    def backward(ctx, grad_output, grad_boundary):
        return ops.interp2x_boundary3d_backward(grad_output.contiguous()), None


class Interp2xBoundary3d(nn.Module):
    def __init__(self, balance_value=0.5):
        super().__init__()
        self.balance_value = balance_value

    def forward(self, input):
        return Interp2xBoundary3dFunction.apply(input, self.balance_value)
