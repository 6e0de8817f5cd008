"""Drop-in for the reference's `GridSamplerMine` extension module
(MCAcc/cuda/GridSamplerMine.cpp:99-103), as imported by MCAcc/grid_sampler_mine.py:6.
Only Bilinear(0) / Border(1) exist, as in the reference's own check (GridSamplerMine.cpp:59-64)."""
from recmv_b200 import ops


def _check_modes(interpolation_mode, padding_mode):
    if interpolation_mode != 0:
        raise RuntimeError("grid_sampler(): only support Bilinear now")
    if padding_mode != 1:
        raise RuntimeError("grid_sampler(): only support Border Padding now")


def forward(input, grid, interpolation_mode, padding_mode):
    _check_modes(interpolation_mode, padding_mode)
    return ops.grid_sample3d_forward(input, grid)


def backward(input, grid, grad_output, interpolation_mode, padding_mode):
    _check_modes(interpolation_mode, padding_mode)
    return ops.grid_sample3d_backward(input, grid, grad_output)


def dbackward(grad_output_input, grad_output_grid, input, grid, grad_output, interpolation_mode,
              padding_mode):
    _check_modes(interpolation_mode, padding_mode)
    return ops.grid_sample3d_dbackward(grad_output_input, grad_output_grid, input, grid, grad_output)
