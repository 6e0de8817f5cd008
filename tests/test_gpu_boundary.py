"""GPU tests of the drop-in boundary: the compat modules under the reference's own module names, the
LBSkinner methods infer_fl.py calls (repose, query_skinning_weights_colors, posedSkeleton), the bone-matrix kernel,
FindSurfacePs on the device, the fp16 operand-range report of the tcgen05 engine, and an SdfRenderer built around
an EXISTING network / skinner pair."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, rel_err
from recmv_b200 import _lib, ops, synth
from recmv_b200.model import LBSkinner, getTmpSdf

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _skinner(extra=None):
    Js, parents, init = synth.skeleton()
    ws = synth.skinning_voxel((17, 33, 21), seed=7)
    return LBSkinner(ws, [-1.1] * 3, [1.1] * 3, Js, parents, init_pose=init, extra_trans=extra,
                     bbox_extend=torch.tensor(synth.BBOX_EXTEND), bbox_center=torch.tensor(synth.BBOX_CENTER)).to(DEV)


def test_compat_modules_under_reference_names():
    """`import FastMinv / MCGpu / GridSamplerMine / interp2x_boundary3d` resolve to this package when
    recmv_b200/compat is first on sys.path, with the reference's call signatures."""
    compat = os.path.join(ROOT, "recmv_b200", "compat")
    saved = {k: sys.modules.pop(k, None) for k in ("FastMinv", "MCGpu", "GridSamplerMine", "interp2x_boundary3d")}
    sys.path.insert(0, compat)
    try:
        import FastMinv
        import GridSamplerMine
        import MCGpu
        import interp2x_boundary3d
        assert FastMinv.__file__.startswith(compat) and MCGpu.__file__.startswith(compat)
        ms = torch.randn((100, 3, 3), generator=synth.generator(1)).to(DEV) + 2 * torch.eye(3, device=DEV)
        invs, checks = FastMinv.Fast3x3Minv(ms)                      # FastMinv/M3x3Inv.cpp:12-37
        assert checks.dtype == torch.bool and (invs @ ms - torch.eye(3, device=DEV)).abs().max() < 1e-4
        outs = FastMinv.Fast3x3Minv_backward(torch.ones_like(invs), invs)
        assert outs.shape == invs.shape
        MCGpu.mc_init(0)
        sdf = synth.sphere_sdf_grid(33, num=2, seed=3, device=DEV)
        verts, faces = MCGpu.mc_gpu(sdf, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0)   # MCGpu.cpp:20-56 positional form
        assert verts.dtype == torch.float32 and faces.dtype == torch.int64 and faces.shape[0] > 0
        assert MCGpu.mc_gpu(sdf.double()) == []
        inp = torch.randn((1, 5, 9, 9, 9), device=DEV)
        grid = (torch.rand((1, 1, 1, 50, 3), device=DEV) - 0.5) * 2.2
        out = GridSamplerMine.forward(inp, grid, 0, 1)
        ref = torch.nn.functional.grid_sample(inp, grid, mode="bilinear", padding_mode="border", align_corners=False)
        assert (out - ref).abs().max() < 1e-5
        gi, gg = GridSamplerMine.backward(inp, grid, torch.ones_like(out), 0, 1)
        d0, d1, d2 = GridSamplerMine.dbackward(torch.zeros_like(inp), torch.ones_like(grid), inp, grid,
                                               torch.ones_like(out), 0, 1)
        assert gi.shape == inp.shape and gg.shape == grid.shape and d2.shape == out.shape
        with pytest.raises(RuntimeError):
            GridSamplerMine.forward(inp, grid, 1, 1)                   # only Bilinear, GridSamplerMine.cpp:59-64
        o, b = interp2x_boundary3d.forward(torch.randn((1, 1, 5, 6, 7), device=DEV), 0.0)
        assert o.shape == (1, 1, 9, 11, 13) and b.dtype == torch.bool
        assert interp2x_boundary3d.backward(torch.ones_like(o)).shape == (1, 1, 5, 6, 7)
    finally:
        sys.path.remove(compat)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_bone_matrix_kernel_matches_torch_chain():
    for init in (True, False):
        sk = _skinner()
        if not init:
            sk.init_pose = None
        poses, _ = synth.poses_trans(5, seed=3)
        poses[0] = 0.0                                     # zero rotation: the |theta + 1e-8| branch
        poses = poses.to(DEV)
        A = sk.bone_matrices(poses)                        # kernel (no grad)
        pg = poses.clone().requires_grad_(True)
        A_ref = sk.bone_matrices(pg)                       # torch chain (autograd)
        assert A_ref.requires_grad and not A.requires_grad
        assert (A - A_ref).abs().max() < 2e-6
        J = sk.posedSkeleton([poses, torch.zeros((5, 3), device=DEV)])
        Jr = sk.posedSkeleton([pg, torch.zeros((5, 3), device=DEV)])
        assert (J - Jr).abs().max() < 2e-6


def test_repose_and_weight_colours():
    extra = torch.tensor([[0.1, -0.2, 0.05]])
    sk = _skinner(extra)
    g = load_golden("lbs.npz")
    poses, trans = torch.from_numpy(g["poses"]).to(DEV), torch.from_numpy(g["trans"]).to(DEV)
    ps, bi = torch.from_numpy(g["ps"]).to(DEV), torch.from_numpy(g["batch_inds"]).to(DEV)
    with torch.no_grad():
        fwd = sk(ps, [poses, trans], bi)
        rep = sk.repose(ps, [poses, trans], bi)            # model/Deformer.py:446-531: no extra translation
    assert rel_err(rep, g["out_list"], 1e-2) < 1e-4
    assert (fwd - rep - extra.to(DEV)).abs().max() < 1e-6
    rep_b = sk.repose(ps.view(3, 1000, 3), [poses, trans], None)
    assert rel_err(rep_b, g["out_batch"], 1e-2) < 1e-4
    # query_skinning_weights_colors (model/Deformer.py:331-340): weights x fixed joint colours, on the CPU in float64
    cols = sk.query_skinning_weights_colors(ps)
    nps = ((ps - sk.bbox_center.view(1, 3)) / sk.bbox_extend * 2).view(1, 1, 1, -1, 3)
    w = torch.nn.functional.grid_sample(sk.ws, nps, mode="bilinear", padding_mode="border",
                                        align_corners=False).view(24, -1).t()
    assert cols.shape == (3000, 3) and cols.dtype == torch.float64 and not cols.is_cuda
    ref = (w.cpu().double()[:, :, None] * LBSkinner._JOINT_COLORS[None]).sum(1)
    assert (cols - ref).abs().max() < 1e-5
    assert cols.min() >= 0 and cols.max() <= 1.0 + 1e-5
    assert abs(float(LBSkinner._JOINT_COLORS[1, 2]) - 180 / 255) < 1e-12    # joint 1 = 'blue' = Paired[1]


def test_findsurfaceps_on_device_matches_reference_golden():
    from recmv_b200.utils import FindSurfacePs

    class Frags:
        def __init__(self, p2f, bary):
            self.pix_to_face, self.bary_coords = p2f, bary
    g = load_golden("findsurface.npz")
    t = {k: torch.from_numpy(v).to(DEV) for k, v in g.items()}
    b, r, c, pts, fi = FindSurfacePs(t["verts"], t["faces"], Frags(t["pix_to_face"], t["bary"]))
    assert b.is_cuda and b.dtype == torch.int64
    assert np.array_equal(b.cpu().numpy(), g["batch"]) and np.array_equal(r.cpu().numpy(), g["row"])
    assert np.array_equal(c.cpu().numpy(), g["col"]) and np.array_equal(fi.cpu().numpy(), g["finds"])
    assert np.allclose(pts.cpu().numpy(), g["pts"], atol=1e-6)
    e = FindSurfacePs(t["verts"], t["faces"], Frags(torch.full((1, 4, 4, 1), -1, device=DEV),
                                                   torch.rand((1, 4, 4, 1, 3), device=DEV)))
    assert all(x.numel() == 0 for x in e[:3]) and e[3].shape == (0, 3)


def test_fragment_decode_with_mask_and_view_rays():
    """(f3): FindSurfacePs + the gt-mask selection and view_rays of sample_train_ray in one device pass, against the
    composition of the reference formulas in torch (CameraMine.py:146-167, OptimGarmentNetwork.py:1006-1011)."""
    from recmv_b200.utils import FindSurfacePs, FindSurfacePsRays

    class Frags:
        def __init__(self, p2f, bary):
            self.pix_to_face, self.bary_coords = p2f, bary
    g = synth.generator(123)
    V, Fc, N, H, W, K = 300, 500, 3, 64, 48, 2
    verts = torch.randn((V, 3), generator=g).to(DEV)
    faces = torch.randint(0, V, (Fc, 3), generator=g).to(DEV)
    p2f = torch.randint(-1, N * Fc, (N, H, W, K), generator=g)
    p2f[torch.rand((N, H, W, K), generator=g) < 0.5] = -1
    bary = torch.rand((N, H, W, K, 3), generator=g)
    bary[torch.rand((N, H, W, K), generator=g) < 0.2] *= -1.0
    mask = (torch.rand((N, H, W), generator=g) > 0.3).float().to(DEV)
    fr = Frags(p2f.to(DEV), bary.to(DEV))
    fx, fy, px, py = 1.2 * W, 1.1 * W, W / 2 - 0.3, H / 2 + 0.2
    R = torch.linalg.qr(torch.randn((3, 3), generator=g))[0]
    b, r, c, pts, fi = FindSurfacePs(verts, faces, fr)                      # device pass, no mask
    sel = mask[b, r, c] > 0.
    ps = torch.stack([c[sel].float(), r[sel].float(), torch.ones_like(c[sel]).float()], 1)
    rays = torch.zeros_like(ps)
    rays[:, 0] = -ps[:, 0] / fx + ps[:, 2] * px / fx
    rays[:, 1] = -ps[:, 1] / fy + ps[:, 2] * py / fy
    rays[:, 2] = ps[:, 2]
    rays = (rays / torch.norm(rays, p=2, dim=1, keepdim=True)).matmul(R.to(DEV).transpose(0, 1))
    b2, r2, c2, pts2, fi2, rays2 = FindSurfacePsRays(verts, faces, fr, (fx, fy, px, py, R), mask)
    assert torch.equal(b2, b[sel]) and torch.equal(r2, r[sel]) and torch.equal(c2, c[sel]) and torch.equal(fi2, fi[sel])
    assert torch.equal(pts2, pts[sel]) and (rays2 - rays).abs().max() < 1e-6
    assert (rays2.norm(dim=1) - 1).abs().max() < 1e-5 and b2.numel() > 1000


def test_fp16_operand_range_is_reported_not_silent():
    """|a| >= 1023.5 or |w| >= 63.97 leave the fp16 range of the tcgen05 operands.  The kernels saturate
    and raise status code 2; recmv_check_async_errors reports it (and later launches are refused until cleared)."""
    import recmv_b200.model as M
    torch.manual_seed(2)
    rn = M.RenderingNetwork_view_norm(256, d_in=9, d_out=3, dims=[512] * 4, mode="idr", weight_norm=True,
                                      multires_v=4, multires_n=0).to(DEV)
    P = 256
    pts = torch.zeros((P, 3), device=DEV)
    unit = torch.nn.functional.normalize(torch.randn((P, 3), device=DEV), dim=1)
    ratio = {"renderRatio": None}
    with torch.no_grad():
        ok = rn(pts, unit, unit, torch.randn((P, 256), device=DEV), ratio)
        torch.cuda.synchronize()
        ops.check_async_errors()
        assert torch.isfinite(ok).all()
        bad = rn(pts, unit, unit, torch.full((P, 256), 5000.0, device=DEV), ratio)   # 64 * 5000 > 65504
        torch.cuda.synchronize()
        assert torch.isfinite(bad).all()                     # saturated, not NaN
        with pytest.raises(_lib.RecmvError, match="range"):
            ops.check_async_errors(clear=True)
        ops.check_async_errors()                             # cleared
        # weights out of range are caught at pack time
        with torch.no_grad():
            rn.lin1.weight_g.mul_(1e4)
        rn(pts, unit, unit, torch.randn((P, 256), device=DEV), ratio)
        torch.cuda.synchronize()
        with pytest.raises(_lib.RecmvError, match="range"):
            ops.check_async_errors(clear=True)


def test_renderer_takes_an_existing_scene():
    from recmv_b200.render import SdfRenderer
    torch.manual_seed(0)
    net = getTmpSdf(DEV, 6, 0.6, 256)
    sk = _skinner()
    ren = SdfRenderer(DEV, sdf_net=net, skinner=sk, samples=16)
    assert ren.skinner is sk and ren.sdf_net is net
    poses, trans = synth.poses_trans(1, seed=11)
    A, t = ren.bone_matrices(poses.to(DEV), trans.to(DEV))
    dirs = synth.pinhole_rays(16, 16, device=DEV)
    s0 = ren.render(dirs, A, t)[0].clone()
    with torch.no_grad():                                    # an optimizer step between renders must be picked up
        net.lin8.bias.add_(0.05)
    s1 = ren.render(dirs, A, t)[0]
    fin = s0 < 1e9
    assert fin.any() and ((s1 - s0)[fin] - 0.05).abs().max() < 1e-5
