"""CPU checks of the host-side pieces of the second-order path (recmv_b200/second_order.py, utils) and of the
consistency of its fixtures (tests/golden/make_golden_f4.py).  The GEMM path itself is GPU-only (test_gpu_second_order.py)."""
import numpy as np
import torch

from conftest import load_golden
from recmv_b200 import ops, second_order, utils


def test_pe_vjp_is_the_transpose_jacobian_of_the_positional_encoding():
    g = torch.Generator().manual_seed(3)
    x = (torch.rand((50, 3), generator=g, dtype=torch.float64) - 0.5).requires_grad_(True)
    u = torch.randn((50, 39), generator=g, dtype=torch.float64, requires_grad=True)
    pe_w = [0.3 + 0.05 * k for k in range(12)]
    ref = torch.autograd.grad((ops._pe_torch(x, pe_w, 6) * u).sum(), x, create_graph=True)[0]
    ours = second_order.pe_vjp_torch(x, u, pe_w, 6)
    assert float((ours - ref).abs().max()) < 1e-13
    # and its derivatives (what autograd differentiates in the eikonal backward)
    w = torch.randn((50, 3), generator=g, dtype=torch.float64)
    for wrt in (x, u):
        a = torch.autograd.grad((ours * w).sum(), wrt, retain_graph=True)[0]
        b = torch.autograd.grad((ref * w).sum(), wrt, retain_graph=True)[0]
        assert float((a - b).abs().max()) < 1e-12


def test_input_grad_only_nests_and_unwinds():
    assert ops._INPUT_GRAD_ONLY[0] == 0 and not ops._inputs_only(True) and ops._inputs_only(False)
    with ops.input_grad_only():
        with ops.input_grad_only():
            assert ops._inputs_only(True)
        assert ops._inputs_only(True)
    try:
        with ops.input_grad_only():
            raise ValueError
    except ValueError:
        pass
    assert ops._INPUT_GRAD_ONLY[0] == 0 and not ops._inputs_only(True)


def test_regulariser_fixture_is_self_consistent_and_robust_error_matches():
    g = load_golden("def_regu.npz")
    s = np.linalg.svd(g["J_f64"], compute_uv=False)                       # the SVD oracle: LAPACK in float64
    assert np.abs(s - g["s_f64"]).max() < 1e-12
    sl = torch.log(torch.from_numpy(g["s_f64"]))
    loss = utils.GMRobustError((sl * sl).sum(1), float(g["c"]), True).mean()
    assert abs(float(loss) - float(g["loss_f64"])) < 1e-15
    x = torch.tensor([0.0, 0.3, 2.0], dtype=torch.float64)
    c = 0.2
    assert torch.allclose(utils.GMRobustError(x, c, True), 2 * x / c ** 2 / (x / c ** 2 + 4))
    assert torch.allclose(utils.GMRobustError(x, c, False), 2 * x * x / c ** 2 / (x * x / c ** 2 + 4))


def test_eikonal_fixture_is_self_consistent():
    for tag in ("geo", "trained"):
        g = load_golden(f"sdf_eikonal_{tag}.npz")
        loss = float(((g["gnorm_f64"] - 1.0) ** 2).mean())
        assert abs(loss - float(g["loss_f64"])) < 1e-14
        assert abs(float(g["loss_f32"]) - loss) < 1e-5 * loss
