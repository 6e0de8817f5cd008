import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np
from recmv_b200 import ops, synth, utils
import recmv_b200.model as M
import make_golden as mg
DEV = "cuda:0"
class _Mods:
    getTmpSdf = staticmethod(M.getTmpSdf)
    MLPTranslator, LBSkinner, CompositeDeformer = M.MLPTranslator, M.LBSkinner, M.CompositeDeformer
g = np.load('/root/repo/tests/golden/surface.npz')
t = {k: torch.from_numpy(v).to(DEV) for k, v in g.items()}
sdf, deformer = mg.surface_scene(_Mods, _Mods, device="cpu")
sdf, deformer = sdf.to(DEV), deformer.to(DEV)
torch.manual_seed(2)
rn = M.RenderingNetwork_view_norm(256, d_in=9, d_out=3, dims=[512] * 4, mode="idr", weight_norm=True, multires_v=4, multires_n=0).to(DEV)
ratio = {"sdfRatio": 0.8, "deformerRatio": 0.6, "renderRatio": 0.9}
conds = t["conds"].clone().requires_grad_(True)
target = torch.rand((t["ps"].shape[0], 3), generator=synth.generator(8)).to(DEV)
mods = [sdf, deformer.defs[0], rn]
def merr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))
def step(flags, what):
    for m, f in zip(mods, flags):
        m.train_fused = f
        m.zero_grad(set_to_none=True)
    conds.grad = None
    ps = t["ps"].clone().requires_grad_(True)
    defconds = [conds, [t["poses"], t["trans"]]]
    nrm, ds = utils.compute_deformed_normals(sdf, deformer, ps, defconds, t["batch_inds"], ratio, "train", "body")
    if what == "normals":
        loss = (nrm * target).sum()
    elif what == "ds":
        loss = (ds * target).sum()
    else:
        col = rn(ds, nrm, t["rays"], sdf.rendcond, ratio)
        loss = (col - target).abs().mean()
    loss.backward()
    grads = {f"{type(m).__name__}.{n}": p.grad.clone() for m in mods for n, p in m.named_parameters() if p.grad is not None}
    grads["conds"] = conds.grad.clone() if conds.grad is not None else torch.zeros_like(conds)
    grads["ps"] = ps.grad.clone()
    return grads
for what in ("ds", "normals", "colour"):
    ref = step((False, False, False), what)
    for name, flags in (("sdf", (True, False, False)), ("translator", (False, True, False)), ("rendernet", (False, False, True))):
        gq = step(flags, what)
        errs = sorted(((merr(gq[k], ref[k]), k) for k in ref if k in gq), reverse=True)[:3]
        print(what, "fused:", name, "worst", [(f"{e:.1e}", k) for e, k in errs], "missing", [k for k in ref if k not in gq][:3])
print("---- per-point analysis of ps.grad, colour loss, fused sdf only")
def step2(flags, lossfn):
    for m, f in zip(mods, flags):
        m.train_fused = f
        m.zero_grad(set_to_none=True)
    ps = t["ps"].clone().requires_grad_(True)
    defconds = [conds, [t["poses"], t["trans"]]]
    nrm, ds = utils.compute_deformed_normals(sdf, deformer, ps, defconds, t["batch_inds"], ratio, "train", "body")
    col = rn(ds, nrm, t["rays"], sdf.rendcond, ratio)
    lossfn(col).backward()
    return ps.grad.clone(), col.detach()
for lname, lf in (("L1", lambda c: (c - target).abs().mean()), ("linear", lambda c: (c * target).sum() / c.shape[0])):
    a, ca = step2((False, False, False), lf)
    b, cb = step2((True, False, False), lf)
    per = (a - b).norm(dim=1) / a.norm(dim=1).max()
    print(lname, "P", a.shape[0], "max rel", float(per.max()), "points > 1e-3:", int((per > 1e-3).sum()), "> 1e-4:", int((per > 1e-4).sum()),
          "median", float(per.median()), "col diff", float((ca - cb).abs().max()))
    a2, _ = step2((False, False, False), lf)
    print(lname, "reference run twice: max rel", float(((a - a2).norm(dim=1) / a.norm(dim=1).max()).max()))
