"""CPU: pin the oracle (oracle/oracle_torch.py, oracle/mc_oracle.c) against the reference's own checks
and against the golden vectors produced by the imported reference modules (tests/golden)."""
import numpy as np
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from oracle import mc_oracle
from oracle import oracle_torch as ot
from recmv_b200 import model as M
from recmv_b200 import synth, testing


def _check_params(mod, g):
    c = testing.param_checksums(mod)
    names = [str(n) for n in g["param_names"]]
    assert sorted(c) == names
    got = np.array([c[n] for n in names])
    np.testing.assert_allclose(got, g["param_sums"], rtol=1e-12, atol=1e-9)


def test_gridsample_matches_torch_grid_sample():
    # MCAcc/check_grid_sampler_mine.py:8-9
    g = load_golden("gridsample.npz")
    inp, grid = torch.from_numpy(g["input"]), torch.from_numpy(g["grid"])
    out = ot.grid_sample3d_fwd(inp, grid)
    assert torch.equal(out, F.grid_sample(inp, grid, mode="bilinear", padding_mode="border", align_corners=False)) \
        or (out - torch.from_numpy(g["out"])).abs().max() < 1e-13


def test_gridsample_gradcheck_first_and_second_order():
    # MCAcc/check_grid_sampler_mine.py:10-15 (gradcheck of the function and of its backward)
    gen = synth.generator(9)
    inp = torch.randn((1, 3, 5, 6, 7), generator=gen, dtype=torch.double, requires_grad=True)
    grid = ((torch.rand((1, 1, 1, 6, 3), generator=gen, dtype=torch.double) - 0.5) * 2.2).requires_grad_(True)
    assert torch.autograd.gradcheck(ot.grid_sample3d_fwd, (inp, grid), eps=1e-6, atol=1e-5)
    go = torch.randn((1, 3, 1, 1, 6), generator=gen, dtype=torch.double, requires_grad=True)
    assert torch.autograd.gradcheck(lambda i, g, o: ot.grid_sample3d_bwd(i, g, o), (inp, grid, go), eps=1e-6, atol=1e-5)


def test_minv_inverse_property():
    # FastMinv/check.py:18-20
    ms = torch.randn((10000, 3, 3), generator=synth.generator(0))
    inv, ok = ot.minv3x3_fwd(ms)
    err = (inv[ok] @ ms[ok] - torch.eye(3)).norm(dim=(1, 2))
    assert ok.sum() > 9900 and err.mean() < 1e-4
    sing = torch.zeros(2, 3, 3)
    sing[1] = torch.eye(3) * 0.04  # det = 6.4e-5 < 1e-4
    inv, ok = ot.minv3x3_fwd(sing)
    assert not ok.any() and (inv == 0).all()
    # VJP against autograd of torch.linalg.inv
    m = torch.randn(50, 3, 3, dtype=torch.double, requires_grad=True)
    gr = torch.randn(50, 3, 3, dtype=torch.double)
    ref = torch.autograd.grad(torch.linalg.inv(m), m, gr)[0]
    got = ot.minv3x3_bwd(gr, torch.linalg.inv(m).detach())
    assert (ref - got).abs().max() < 1e-8


def test_sdf_restatement_matches_reference_golden():
    for tag, pseed in (("geo", None), ("trained", 101)):
        g = load_golden(f"sdf_c1_{tag}.npz")
        net = testing.build_sdf(M.getTmpSdf, seed=0, perturb_seed=pseed)
        _check_params(net, g)  # same seeds -> bit-identical parameters as the reference class
        Ws, bs = net.effective_weights()
        Ws = [w.detach() for w in Ws]
        bs = [b.detach() for b in bs]
        x = torch.from_numpy(g["x"])
        for rname, ratio in (("none", None), ("r035", 0.35), ("zero", 0.0)):
            s, f = ot.sdf_mlp(x, Ws, bs, ot.annealing_weights(6, ratio))
            assert rel_err(s[:, 0], g["sdf_" + rname], 1e-2) < 5e-5  # fp32 reassociation noise
            assert rel_err(f[:, ::16], g[f"feat_{rname}_cols"], 1e-2) < 5e-5


def test_translator_and_rendernet_restatements():
    g = load_golden("translator.npz")
    torch.manual_seed(1)
    tr = M.MLPTranslator(128, 6)
    testing.perturb_module(tr, 202, scale=0.5)
    _check_params(tr, g)
    Ws = [getattr(tr, f"lin{l}").weight.detach() for l in range(5)]
    bs = [getattr(tr, f"lin{l}").bias.detach() for l in range(5)]
    p = torch.from_numpy(g["p"])
    cond = torch.from_numpy(g["conds"])[torch.from_numpy(g["batch_inds"])]
    out, off = ot.translator_mlp(p, cond, Ws, bs, ot.annealing_weights(6, 0.6))
    assert rel_err(out, g["out"], 1e-2) < 1e-5 and rel_err(off, g["offset"], 1e-3) < 1e-4

    g = load_golden("rendernet.npz")
    torch.manual_seed(2)
    rn = M.RenderingNetwork_view_norm(256, d_in=9, d_out=3, dims=[512] * 4, mode="idr", weight_norm=True,
                                      multires_v=4, multires_n=0)
    testing.perturb_module(rn, 303)
    _check_params(rn, g)
    Ws = [ot.weight_norm_effective(getattr(rn, f"lin{l}").weight_g, getattr(rn, f"lin{l}").weight_v).detach()
          for l in range(5)]
    bs = [getattr(rn, f"lin{l}").bias.detach() for l in range(5)]
    col = ot.render_mlp(*(torch.from_numpy(g[k]) for k in ("points", "normals", "view_dirs", "feats")),
                        Ws, bs, ot.annealing_weights(4, 0.8))
    assert rel_err(col, g["out"], 1e-2) < 1e-5


def test_lbs_restatement_matches_reference_golden():
    g = load_golden("lbs.npz")
    Js, parents, init = synth.skeleton()
    ws = synth.skinning_voxel((17, 33, 21), seed=7)
    A = ot.bone_matrices(torch.from_numpy(g["poses"]), Js, parents, torch.from_numpy(g["init_pose"]))
    assert (A - torch.from_numpy(g["A"])).abs().max() < 1e-6
    out = ot.lbs_forward(torch.from_numpy(g["ps"]), A, torch.from_numpy(g["trans"]), ws,
                         torch.tensor(synth.BBOX_CENTER), synth.BBOX_EXTEND, torch.from_numpy(g["batch_inds"]))
    assert rel_err(out, g["out_list"], 1e-2) < 5e-5
    assert rel_err(g["out_batch"].reshape(-1, 3)[:1000], g["out_batch"].reshape(-1, 3)[:1000], 1) == 0
    # inverse warp round trip: weights are sampled at different points, so only near-identity poses
    # invert exactly; here just pin the fixture.
    xc, ok = ot.lbs_inverse(torch.from_numpy(g["out_list"]), A, torch.from_numpy(g["trans"]), ws,
                            torch.tensor(synth.BBOX_CENTER), synth.BBOX_EXTEND, torch.from_numpy(g["batch_inds"]))
    assert rel_err(xc, g["inv_xc"], 1e-2) < 5e-5 and bool((ok.numpy() == g["inv_ok"]).all())


def test_mc_oracle_closed_manifold_and_sizes():
    sdf = synth.sphere_sdf_grid(41, num=4, seed=3).numpy()
    v, f = mc_oracle.marching_cubes(sdf, (0.05,) * 3, (-1, -1, -1))
    assert f.min() >= 0 and f.max() == len(v) - 1
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()  # watertight
    # every vertex lies on a lattice edge: two coordinates are lattice-aligned
    lat = (v + 1) / 0.05
    frac = np.abs(lat - np.round(lat))
    assert (np.sort(frac, 1)[:, :2] < 1e-3).all()
    # empty / full grids and a ragged (anisotropic) one
    for arr in (np.ones((5, 6, 7), np.float32), -np.ones((5, 6, 7), np.float32)):
        v0, f0 = mc_oracle.marching_cubes(arr)
        assert len(v0) == 0 and len(f0) == 0
    v1, f1 = mc_oracle.marching_cubes(synth.sphere_sdf_grid((21, 37, 13), num=2, seed=5).numpy())
    assert len(v1) > 0 and f1.max() == len(v1) - 1
