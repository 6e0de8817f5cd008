"""Golden-vector generator, round-2 additions (CONTAINER ONLY: needs /root/reference).

  sdf_c1_{geo,trained}_f64.npz : the reference ImplicitNetwork evaluated in float64 on the C1 inputs (same
      parameters as the fp32 fixtures, cast up) -- the ground truth that tells how far the REFERENCE's own fp32
      result is from exact arithmetic, so the parity tests can assert "no worse than the reference" instead of a
      bar the reference itself misses.
  sdf_bwd_{geo,trained}.npz    : first-order training gradients of the reference class through torch autograd
      (fp32, what train.py:325 `loss.backward()` produces) for L = sum(c_sdf * sdf) + sum(c_feat * feat) with
      seeded cotangents: dL/dx, dL/d bias (full), dL/d weight_g (full), dL/d weight_v as strided samples +
      float64 row/column sums; the same in float64 as ground truth.
  translator_bwd.npz / rendernet_bwd.npz : the same for MLPTranslator and RenderingNetwork_view_norm.
  c2f_grid.npz generator lives in make_golden_c2f.py.

Re-run with:  python tests/golden/make_golden_r2.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
from recmv_b200 import synth, testing  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote", name, os.path.getsize(path) // 1024, "KiB")


def cotangents(P, nfeat, seed):
    """Seeded upstream gradients; per-sample magnitudes spread over two decades like a mean-reduced loss with
    masked terms (the tests regenerate them from the same seed)."""
    g = synth.generator(seed)
    mag = torch.exp(torch.randn((P, 1), generator=g) * 1.0) / P
    c0 = torch.randn((P, 1), generator=g) * mag
    c1 = torch.randn((P, nfeat), generator=g) * mag * 0.1
    return c0, c1


def grad_summary(prefix, named_grads, out):
    """Full vectors for small tensors; for matrices a strided sample + float64 row / column sums."""
    for name, gt in named_grads:
        key = f"{prefix}{name.replace('.', '_')}"
        gt = gt.detach()
        if gt.dim() == 2 and gt.numel() > 4096:
            out[key + "_sample"] = gt[::16, ::8].float()
            out[key + "_rowsum"] = gt.double().sum(1)
            out[key + "_colsum"] = gt.double().sum(0)
            out[key + "_absmax"] = gt.abs().max().double()
        else:
            out[key] = gt.float() if gt.dtype == torch.float32 else gt.double()


def main():
    torch.set_num_threads(8)
    ns = refload.load()
    for tag, pseed in (("geo", None), ("trained", 101)):
        net = testing.build_sdf(ns.network.getTmpSdf, seed=0, perturb_seed=pseed)
        x = torch.rand((4096, 3), generator=synth.generator(1234)) * 1.2 - 0.6
        # ---- fp64 ground truth of the forward (reference class cast to double; same parameters) ------------
        net64 = testing.build_sdf(ns.network.getTmpSdf, seed=0, perturb_seed=pseed).double()
        outs = {}
        for rname, ratio in (("none", None), ("r035", 0.35), ("zero", 0.0)):
            with torch.no_grad():
                y = net64(x.double(), {'sdfRatio': None} if ratio is None else float(ratio))
            outs["sdf_" + rname] = y[:, 0]
            outs[f"feat_{rname}_cols"] = net64.rendcond[:, ::16].clone()
            outs[f"feat_{rname}_rowsum"] = net64.rendcond.sum(1)
        xg = x.double().clone().requires_grad_(True)
        grad64 = torch.autograd.grad(net64(xg, {'sdfRatio': None}).sum(), xg)[0]
        save(f"sdf_c1_{tag}_f64.npz", grad_none=grad64, **outs)

        # ---- first-order training gradients (fp32 = the reference's own numbers; fp64 = truth) -------------
        P = 2048
        c0, c1 = cotangents(P, 256, 77)
        res = {}
        for prec, mod in (("f32", net), ("f64", net64)):
            dt = torch.float32 if prec == "f32" else torch.float64
            mod.zero_grad()
            xi = x[:P].to(dt).clone().requires_grad_(True)
            y = mod(xi, {'sdfRatio': 0.7})
            loss = (y * c0.to(dt)).sum() + (mod.rendcond * c1.to(dt)).sum()
            loss.backward()
            res[f"dx_{prec}"] = xi.grad.clone()
            grad_summary(f"{prec}_", [(n, p.grad) for n, p in sorted(mod.named_parameters())], res)
        save(f"sdf_bwd_{tag}.npz", **res)

    # ---- translator ------------------------------------------------------------------------------------------
    def make_tr():
        torch.manual_seed(1)
        return testing.perturb_module(ns.Deformer.MLPTranslator(128, 6), 202, scale=0.5)
    g = synth.generator(77)
    p = torch.rand((2048, 3), generator=g) * 1.2 - 0.6
    conds = torch.randn((3, 128), generator=g) * 0.1
    binds = torch.randint(0, 3, (2048,), generator=g)
    cot, _ = cotangents(2048, 1, 78)
    cot3 = cot.expand(-1, 3) * torch.randn((2048, 3), generator=synth.generator(79))
    res = {}
    for prec in ("f32", "f64"):
        dt = torch.float32 if prec == "f32" else torch.float64
        mod = make_tr().to(dt)
        pi = p.to(dt).clone().requires_grad_(True)
        ci = conds.to(dt).clone().requires_grad_(True)
        out = mod(pi, ci, binds, ratio={"deformerRatio": 0.6}, offset_type="body")
        (out * cot3.to(dt)).sum().backward()
        res[f"dp_{prec}"] = pi.grad.clone()
        res[f"dconds_{prec}"] = ci.grad.clone()
        grad_summary(f"{prec}_", [(n, q.grad) for n, q in sorted(mod.named_parameters())], res)
    save("translator_bwd.npz", cot=cot3, **res)

    # ---- colour network -----------------------------------------------------------------------------------------
    def make_rn():
        torch.manual_seed(2)
        rn = ns.RenderNet.RenderingNetwork_view_norm(256, d_in=9, d_out=3, dims=[512] * 4, mode="idr",
                                                     weight_norm=True, multires_v=4, multires_n=0)
        return testing.perturb_module(rn, 303)
    pts = torch.rand((1024, 3), generator=g) - 0.5
    nrm = torch.nn.functional.normalize(torch.randn((1024, 3), generator=g), dim=1)
    vd = torch.nn.functional.normalize(torch.randn((1024, 3), generator=g), dim=1)
    feat = torch.randn((1024, 256), generator=g) * 0.3
    cot = torch.randn((1024, 3), generator=synth.generator(80)) / 1024
    res = {}
    for prec in ("f32", "f64"):
        dt = torch.float32 if prec == "f32" else torch.float64
        mod = make_rn().to(dt)
        ins = [t.to(dt).clone().requires_grad_(True) for t in (pts, nrm, vd, feat)]
        col = mod(ins[0], ins[1], ins[2], ins[3], {"renderRatio": 0.8})
        (col * cot.to(dt)).sum().backward()
        for nme, t in zip(("dpoints", "dnormals", "dview", "dfeats"), ins):
            res[f"{nme}_{prec}"] = t.grad.clone() if nme != "dfeats" else t.grad[:, ::8].clone()
        res[f"dfeats_rowsum_{prec}"] = ins[3].grad.double().sum(1)
        grad_summary(f"{prec}_", [(n, q.grad) for n, q in sorted(mod.named_parameters())], res)
    save("rendernet_bwd.npz", cot=cot, points=pts, normals=nrm, view_dirs=vd, feats=feat, **res)


if __name__ == "__main__":
    main()
