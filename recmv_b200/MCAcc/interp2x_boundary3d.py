"""2x-1 upsampler with boundary flags under the reference's names (MCAcc/interp2x_boundary3d.py:9-30): the autograd
function and the small module `Seg3dLossless(use_cuda_impl=True)` instantiates.  Both call the C-ABI kernels
(`recmv_interp2x_boundary3d_{fwd,bwd}`, rounding mode 0 = the reference extension's)."""
import torch
from torch.autograd.function import once_differentiable

from .. import ops


class Interp2xBoundary3dFunction(torch.autograd.Function):
    """(volume [N,C,D,H,W] f32, balance) -> (upsampled [N,C,2D-1,2H-1,2W-1], is_boundary bool of the same shape)."""

    @staticmethod
    def forward(ctx, volume, balance_value):
        if volume.dim() != 5:
            raise RuntimeError("Interp2xBoundary3d expects a [N,C,D,H,W] tensor")
        upsampled, mixed = ops.interp2x_boundary3d_forward(volume.contiguous(), balance_value, 0)
        ctx.mark_non_differentiable(mixed)
        return upsampled, mixed

    @staticmethod
    @once_differentiable
    def backward(ctx, g_upsampled, _g_mixed):
        # the adjoint of the linear upsampling; the flags carry no gradient, nor does the threshold
        return ops.interp2x_boundary3d_backward(g_upsampled.contiguous()), None


class Interp2xBoundary3d(torch.nn.Module):
    def __init__(self, balance_value=0.5):
        super().__init__()
        self.balance_value = balance_value

    def extra_repr(self):
        return f"balance_value={self.balance_value}"

    def forward(self, input):
        return Interp2xBoundary3dFunction.apply(input, self.balance_value)
