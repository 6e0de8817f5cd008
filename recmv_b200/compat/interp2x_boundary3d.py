"""Drop-in for the reference's optional `interp2x_boundary3d` extension module
(MCAcc/cuda/interp2x_boundary3d.cpp:33-36): `forward(input, balance_value) -> [output, is_boundary]`,
`backward(grad_output) -> grad_input`, used by MCAcc/interp2x_boundary3d.py when Seg3dLossless is built with
use_cuda_impl=True."""
from recmv_b200.ops import interp2x_boundary3d_backward as backward  # noqa: F401
from recmv_b200.ops import interp2x_boundary3d_forward as forward  # noqa: F401
