"""Golden-vector generator for the second-order terms of the training step (CONTAINER ONLY: needs /root/reference).

  sdf_eikonal_{geo,trained}.npz : the eikonal term exactly as the reference writes it
      (engineer/networks/OptimGarmentNetwork.py:1108-1118): pred = net(x, ratio); grad = net.gradient(x, pred)
      [model/network.py:121-133, create_graph=True]; loss = ((|grad| - 1)^2).mean(); loss.backward().
      Stored: loss, |grad| per point, dL/dx, and every parameter gradient (bias / weight_g full, weight_v as a strided
      sample + float64 row / column sums) -- in float32 (the reference's own numbers) and float64 (ground truth).
  def_regu.npz : the deformation regulariser (OptimGarmentNetwork.py:1135-1154) of the reference MLPTranslator:
      defVs = translator(p [N,V,3], cond); J = utils.compute_Jacobian(p, defVs, True, True); s = svd(J.cpu()).S;
      loss = GMRobustError(sum(log(s)^2), c, True).mean(); loss.backward().  Stored like the above, plus J and s.

Re-run with:  python tests/golden/make_golden_f4.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_r2 import grad_summary, save  # noqa: E402
from oracle import refload  # noqa: E402
from recmv_b200 import synth, testing  # noqa: E402

EIK_P = 1536
EIK_RATIO = 0.7
REGU_N, REGU_V, REGU_C = 2, 640, 0.2


def main():
    torch.set_num_threads(8)
    ns = refload.load()
    for tag, pseed in (("geo", None), ("trained", 101)):
        x = torch.rand((EIK_P, 3), generator=synth.generator(4321)) * 1.4 - 0.7
        res = {"x": x}
        for prec in ("f32", "f64"):
            dt = torch.float32 if prec == "f32" else torch.float64
            net = testing.build_sdf(ns.network.getTmpSdf, seed=0, perturb_seed=pseed).to(dt)
            xi = x.to(dt).clone().requires_grad_(True)
            pred = net(xi, {'sdfRatio': EIK_RATIO})
            grad = net.gradient(xi, pred)
            loss = ((grad.norm(2, dim=-1) - 1) ** 2).mean()
            loss.backward()
            res[f"loss_{prec}"] = loss.detach().double()
            res[f"gnorm_{prec}"] = grad.detach().norm(2, dim=-1)
            res[f"dx_{prec}"] = xi.grad.clone()
            grad_summary(f"{prec}_", [(n, p.grad) for n, p in sorted(net.named_parameters()) if p.grad is not None], res)
        save(f"sdf_eikonal_{tag}.npz", **res)

    def make_tr():
        torch.manual_seed(1)
        return testing.perturb_module(ns.Deformer.MLPTranslator(128, 6), 202, scale=0.5)
    g = synth.generator(99)
    p = torch.rand((REGU_N, REGU_V, 3), generator=g) * 1.2 - 0.6
    conds = torch.randn((REGU_N, 128), generator=g) * 0.1
    res = {"p": p, "conds": conds, "c": np.float64(REGU_C)}
    for prec in ("f32", "f64"):
        dt = torch.float32 if prec == "f32" else torch.float64
        mod = make_tr().to(dt)
        pi = p.to(dt).clone().requires_grad_(True)
        ci = conds.to(dt).clone().requires_grad_(True)
        defVs = mod(pi, ci, ratio={"deformerRatio": 0.6}, offset_type="body")
        J = ns.utils_utils.compute_Jacobian(pi, defVs, True, True)
        _, s, _ = torch.svd(J.cpu())
        sl = torch.log(s)
        loss = ns.utils_utils.GMRobustError((sl * sl).sum(1), REGU_C, True).mean()
        loss.backward()
        res[f"loss_{prec}"] = loss.detach().double()
        res[f"J_{prec}"] = J.detach()
        res[f"s_{prec}"] = s.detach()
        res[f"dconds_{prec}"] = ci.grad.clone()
        res[f"dp_{prec}"] = pi.grad.clone()
        grad_summary(f"{prec}_", [(n, q.grad) for n, q in sorted(mod.named_parameters()) if q.grad is not None], res)
    save("def_regu.npz", **res)


if __name__ == "__main__":
    main()
