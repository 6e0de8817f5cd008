"""Sweep the accumulator gain of the tcgen05 modes against the reference goldens (C1: geometric-init and
'trained-like' SDF networks, three annealing ratios; translator; colour net) and print the error metrics."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from recmv_b200 import _lib, testing
from recmv_b200.model import getTmpSdf

dev = "cuda:0"
lib = _lib.load()


def norm_err(a, b):
    b = torch.as_tensor(b).to(a.device).float()
    return float((a.float() - b).abs().max() / b.pow(2).mean().sqrt())


def rel_err(a, b, floor):
    b = torch.as_tensor(b).to(a.device).float()
    return float(((a.float() - b).abs() / b.abs().clamp_min(floor)).max())


nets = {t: testing.build_sdf(getTmpSdf, seed=0, perturb_seed=None if t == "geo" else 101).to(dev) for t in ("geo", "trained")}
gold = {t: np.load(os.path.join(ROOT, "tests", "golden", f"sdf_c1_{t}.npz")) for t in nets}
mode = _lib.MLP_TC_F16X3 if len(sys.argv) < 2 or sys.argv[1] == "tc3" else _lib.MLP_TC_F16X1
unit = 4 * 2.0 ** -24      # one 64-wide K block = 4 accumulating MMAs x mean truncation 2^-24
print("alpha (gain per K block / (4 * 2^-24)) | worst norm(sdf) norm(feat) elementwise(sdf) elementwise(feat) rowsum | mean signed feat err")
for alpha in (0.0, 0.5, 0.8, 0.9, 1.0, 1.1, 1.2, 1.5, 2.0):
    assert lib.recmv_tc_set_acc_gain(mode, alpha * unit) == 0
    worst = [0, 0, 0, 0, 0]
    signed = []
    for t, net in nets.items():
        net.mlp_mode = mode
        g = gold[t]
        x = torch.from_numpy(g["x"]).to(dev)
        for rname, ratio in (("none", None), ("r035", 0.35), ("zero", 0.0)):
            with torch.no_grad():
                y = net(x, ratio)
            f = net.rendcond[:, ::16]
            m = [norm_err(y[:, 0], g["sdf_" + rname]), norm_err(f, g[f"feat_{rname}_cols"]),
                 rel_err(y[:, 0], g["sdf_" + rname], 1e-2), rel_err(f, g[f"feat_{rname}_cols"], 1e-2),
                 float((net.rendcond.double().sum(1).cpu() - torch.from_numpy(g[f"feat_{rname}_rowsum"])).abs().max())]
            worst = [max(a, b) for a, b in zip(worst, m)]
            signed.append(float((f.cpu() - torch.from_numpy(g[f"feat_{rname}_cols"])).mean()))
    print(f"{alpha:5.2f} | {worst[0]:.2e} {worst[1]:.2e} {worst[2]:.2e} {worst[3]:.2e} {worst[4]:.2e} | {np.mean(signed):+.2e}")
lib.recmv_tc_set_acc_gain(mode, unit)
