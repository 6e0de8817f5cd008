"""Golden-vector generator (CONTAINER ONLY: needs /root/reference).

Imports the reference's own, unmodified Python modules through oracle/refload.py, runs them on CPU on
seeded inputs and writes small fixtures next to this file.  The GPU box has no /root/reference: the
`-m gpu` tests read these fixtures.  Re-run with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_torch as ot  # noqa: E402
from oracle import refload  # noqa: E402
from recmv_b200 import synth, testing  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote", name, os.path.getsize(path) // 1024, "KiB")


def cks(mod):
    c = testing.param_checksums(mod)
    names = sorted(c)
    return np.array(names), np.array([c[n] for n in names], dtype=np.float64)


def main():
    torch.set_num_threads(8)
    ns = refload.load()

    # ---- C1: SDF MLP on 4096 canonical points, geometric init and trained-like ---------------
    for tag, pseed in (("geo", None), ("trained", 101)):
        net = testing.build_sdf(ns.network.getTmpSdf, seed=0, perturb_seed=pseed)
        g = synth.generator(1234)
        x = torch.rand((4096, 3), generator=g) * 1.2 - 0.6
        outs = {}
        for rname, ratio in (("none", None), ("r035", 0.35), ("zero", 0.0)):
            with torch.no_grad():
                y = net(x, {'sdfRatio': None} if ratio is None else float(ratio))
            outs["sdf_" + rname] = y[:, 0]
            outs["feat_" + rname] = net.rendcond.clone()
        # input gradient (autograd through the reference graph) for ratio None
        xg = x.clone().requires_grad_(True)
        y = net(xg, {'sdfRatio': None})
        grad = torch.autograd.grad(y.sum(), xg)[0]
        names, sums = cks(net)
        # features are stored as 16 strided columns + row sums to keep the fixture small
        small = {k: v for k, v in outs.items() if k.startswith("sdf_")}
        for k, v in outs.items():
            if k.startswith("feat_"):
                small[k + "_cols"] = v[:, ::16]
                small[k + "_rowsum"] = v.double().sum(1)
        save(f"sdf_c1_{tag}.npz", x=x, grad_none=grad, param_names=names, param_sums=sums, **small)

    # ---- translator + colour MLP ----------------------------------------------------------------
    torch.manual_seed(1)
    tr = ns.Deformer.MLPTranslator(128, 6)
    testing.perturb_module(tr, 202, scale=0.5)  # last layer is ~0 at init: make offsets visible
    g = synth.generator(77)
    p = torch.rand((2048, 3), generator=g) * 1.2 - 0.6
    conds = torch.randn((3, 128), generator=g) * 0.1
    binds = torch.randint(0, 3, (2048,), generator=g)
    with torch.no_grad():
        out = tr(p, conds, binds, ratio={"deformerRatio": 0.6}, offset_type="body")
    names, sums = cks(tr)
    save("translator.npz", p=p, conds=conds, batch_inds=binds, out=out, offset=tr.offset["body"],
         param_names=names, param_sums=sums)

    torch.manual_seed(2)
    rn = ns.RenderNet.RenderingNetwork_view_norm(256, d_in=9, d_out=3, dims=[512] * 4, mode="idr",
                                                 weight_norm=True, multires_v=4, multires_n=0)
    testing.perturb_module(rn, 303)
    pts = torch.rand((1024, 3), generator=g) - 0.5
    nrm = torch.nn.functional.normalize(torch.randn((1024, 3), generator=g), dim=1)
    vd = torch.nn.functional.normalize(torch.randn((1024, 3), generator=g), dim=1)
    feat = torch.randn((1024, 256), generator=g) * 0.3
    with torch.no_grad():
        col = rn(pts, nrm, vd, feat, {"renderRatio": 0.8})
    names, sums = cks(rn)
    save("rendernet.npz", points=pts, normals=nrm, view_dirs=vd, feats=feat, out=col,
         param_names=names, param_sums=sums)

    # ---- LBSkinner.forward (reference class + shims), both call forms ------------------------------
    Js, parents, init = synth.skeleton()
    ws = synth.skinning_voxel((17, 33, 21), seed=7)
    sk = ns.Deformer.LBSkinner(ws, [-1.1] * 3, [1.1] * 3, Js, parents, init_pose=init,
                               bbox_extend=torch.tensor(synth.BBOX_EXTEND),
                               bbox_center=torch.tensor(synth.BBOX_CENTER))
    poses, trans = synth.poses_trans(3, seed=11)
    ps = (torch.rand((3000, 3), generator=g) - 0.5) * 2.6  # ~8 % outside the voxel -> border clamp
    bi = torch.randint(0, 3, (3000,), generator=g)
    with torch.no_grad():
        o_list = sk(ps, [poses, trans], bi)
        o_batch = sk(ps.view(3, 1000, 3), [poses, trans], None)
        A = ot.bone_matrices(poses, Js, parents, sk.init_pose)
        inv_xc, inv_ok = ot.lbs_inverse(o_list, A, trans, ws, torch.tensor(synth.BBOX_CENTER),
                                        synth.BBOX_EXTEND, bi)
    save("lbs.npz", ps=ps, batch_inds=bi, poses=poses, trans=trans, out_list=o_list, out_batch=o_batch,
         A=A, init_pose=sk.init_pose, inv_xc=inv_xc, inv_ok=inv_ok)

    # ---- grid sampler: the reference's own check (MCAcc/check_grid_sampler_mine.py) ---------------
    gg = synth.generator(5)
    inp = torch.randn((1, 5, 15, 15, 15), generator=gg, dtype=torch.double)
    grid = (torch.rand((1, 1, 1, 64, 3), generator=gg, dtype=torch.double) - 0.5) * 2.2
    ref = torch.nn.functional.grid_sample(inp, grid, mode="bilinear", padding_mode="border",
                                          align_corners=False)
    go = torch.randn(ref.shape, generator=gg, dtype=torch.double)
    gi, ggr = ot.grid_sample3d_bwd(inp, grid, go)
    ggi = torch.randn(inp.shape, generator=gg, dtype=torch.double)
    ggg = torch.randn(grid.shape, generator=gg, dtype=torch.double)
    d0, d1, d2 = ot.grid_sample3d_bwd2(ggi, ggg, inp, grid, go)
    save("gridsample.npz", input=inp, grid=grid, out=ref, grad_out=go, grad_input=gi.detach(),
         grad_grid=ggr.detach(), gg_input=ggi, gg_grid=ggg, d_input=d0, d_grid=d1, d_gout=d2)

    surface_goldens(ns)


class Frags:
    def __init__(self, pix_to_face, bary_coords):
        self.pix_to_face, self.bary_coords = pix_to_face, bary_coords


def surface_scene(mod_net, mod_def, device="cpu"):
    """Shared by the generator (reference classes) and the GPU tests (this package's classes)."""
    sdf = testing.build_sdf(mod_net.getTmpSdf, seed=0, perturb_seed=101, device=device)
    torch.manual_seed(1)
    tr = mod_def.MLPTranslator(128, 6)
    testing.perturb_module(tr, 202, scale=0.5)
    Js, parents, init = synth.skeleton()
    ws = synth.skinning_voxel((17, 33, 21), seed=7)
    sk = mod_def.LBSkinner(ws, [-1.1] * 3, [1.1] * 3, Js, parents, init_pose=init,
                           bbox_extend=torch.tensor(synth.BBOX_EXTEND), bbox_center=torch.tensor(synth.BBOX_CENTER))
    deformer = mod_def.CompositeDeformer([tr, sk]).to(device)
    return sdf, deformer


def surface_inputs(sdf, deformer):
    """Seeds as the real pipeline produces them: points ON the zero level set (bisection along a radial
    line), rays through their deformed positions, then a 2e-3 perturbation (marching-cubes / rasteriser
    discretisation error) -- the solve only has to polish."""
    g = synth.generator(4242)
    n = 600
    dirs = torch.nn.functional.normalize(torch.randn((n, 3), generator=g), dim=1)
    poses, trans = synth.poses_trans(2, seed=11)
    conds = torch.randn((2, 128), generator=g) * 0.1
    binds = torch.randint(0, 2, (n,), generator=g)
    cam = torch.tensor(synth.CAM_POS)
    ratio = {"sdfRatio": 0.8, "deformerRatio": 0.6, "renderRatio": 0.9}
    lo, hi = torch.full((n, 1), 0.3), torch.full((n, 1), 0.9)
    with torch.no_grad():
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            neg = sdf(dirs * mid, ratio) < 0
            lo, hi = torch.where(neg, mid, lo), torch.where(neg, hi, mid)
        ps_true = dirs * (0.5 * (lo + hi))
        d0 = deformer(ps_true, [conds, [poses, trans]], binds, ratio=ratio, offset_type="body")
    rays = torch.nn.functional.normalize(d0 - cam.view(1, 3), dim=1)
    seeds = ps_true + torch.randn((n, 3), generator=g) * 2e-3
    return ps_true, seeds, rays, poses, trans, conds, binds, cam, ratio


def surface_goldens(ns):
    sdf, deformer = surface_scene(ns.network, ns.Deformer)
    p0, seeds, rays, poses, trans, conds, binds, cam, ratio = surface_inputs(sdf, deformer)
    defconds = [conds, [poses, trans]]
    fsp = ns.FindSurfacePs
    ps, ok = fsp.OptimizeGarmentSurfaceSinlge(cam, rays, seeds.clone(), binds, sdf, ratio, deformer, defconds,
                                              dthreshold=1.e-4, athreshold=0.05, w1=3.05, w2=1., times=10,
                                              offset_type="body")
    ps1, ok1 = fsp.OptimizeSurfacePs(cam, rays, seeds.clone(), binds, sdf, ratio,
                                     lambda p, c, i, **kw: deformer(p, c, i, ratio=kw["ratio"], offset_type="body"),
                                     defconds, dthreshold=1.e-4, athreshold=0.05, times=1)
    # cardinal rays / deformed normals at the solved points (test phase: first order only)
    uu = ns.utils_utils
    pts = ps.detach().clone().requires_grad_(True)
    crays, ds = uu.compute_cardinal_rays(deformer, pts, rays, defconds, binds, ratio, "test", offset_type="body")
    pts2 = ps.detach().clone().requires_grad_(True)
    nrm, ds2 = uu.compute_deformed_normals(sdf, deformer, pts2, defconds, binds, ratio, "test", "body")
    with torch.no_grad():
        sdf_at = sdf(ps, ratio)
    save("surface.npz", ps_true=p0, seeds=seeds, rays=rays, poses=poses, trans=trans, conds=conds, batch_inds=binds,
         ps=ps, ok=ok, ps_1it=ps1, ok_1it=ok1, crays=crays.detach(), ds=ds.detach(),
         normals=nrm.detach(), sdf_at=sdf_at[:, 0])
    print("surface solve: converged", int(ok.sum()), "of", ok.numel(), "| after 1 it:", int(ok1.sum()))

    # FindSurfacePs on synthetic fragments (K = 2 so the first-valid-k logic is exercised)
    g = synth.generator(99)
    V, Fc, N, H, W, K = 50, 80, 2, 12, 10, 2
    verts = torch.randn((V, 3), generator=g)
    faces = torch.randint(0, V, (Fc, 3), generator=g)
    p2f = torch.randint(-1, N * Fc, (N, H, W, K), generator=g)
    p2f[torch.rand((N, H, W, K), generator=g) < 0.4] = -1
    bary = torch.rand((N, H, W, K, 3), generator=g)
    bary[torch.rand((N, H, W, K), generator=g) < 0.2] *= -1.0
    b, r, c, pts0, fi = fsp.FindSurfacePs(verts, faces, Frags(p2f, bary))
    save("findsurface.npz", verts=verts, faces=faces, pix_to_face=p2f, bary=bary, batch=b, row=r, col=c, pts=pts0,
         finds=fi)


if __name__ == "__main__":
    main()
