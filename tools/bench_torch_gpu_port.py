"""Informative only (profiles/): the reference-style PyTorch path -- separate torch ops, fp32 cuBLAS GEMMs, every
activation through HBM -- for the same workload ON THE SAME B200, i.e. the oracle's restatement moved to cuda:0
("the reference path on the same B200 ... the reference single-GPU rays/s the >= 10x target is measured
against").  Round 2: the skinning-voxel sample and the 3x3 inverse run on the REFERENCE'S OWN CUDA kernels
(GridSamplerMine.forward, FastMinv.Fast3x3Minv: built unmodified for sm_100a by oracle/build_ref.py into oracle/_ref);
the MLP is the reference module's op sequence (torch linears + softplus, cuBLAS fp32).  One JSON line per matmul mode."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from oracle import oracle_torch as ot  # noqa: E402
from recmv_b200 import synth  # noqa: E402
from recmv_b200.model import LBSkinner, getTmpSdf  # noqa: E402

dev = torch.device("cuda", 0)
H = W = 512
S = 64
torch.manual_seed(0)
net = getTmpSdf("cpu", 6, 0.6, 256)
Ws, bs = net.effective_weights()
Ws, bs = [w.detach().to(dev) for w in Ws], [b.detach().to(dev) for b in bs]
Js, parents, init = synth.skeleton()
ws = synth.skinning_voxel((65, 225, 129), seed=7)
sk = LBSkinner(ws, [-1.1] * 3, [1.1] * 3, Js, parents, init_pose=init, bbox_extend=torch.tensor(synth.BBOX_EXTEND),
               bbox_center=torch.tensor(synth.BBOX_CENTER))
poses, trans = synth.poses_trans(1, seed=11)
A = ot.bone_matrices(poses, Js, parents, sk.init_pose).to(dev)
trans, ws = trans.to(dev), ws.to(dev)
cam = torch.tensor(synth.CAM_POS, device=dev)
center = torch.tensor(synth.BBOX_CENTER, device=dev)
tk = (synth.T_NEAR + (torch.arange(S, dtype=torch.float32) + 0.5) * (synth.T_FAR - synth.T_NEAR) / S).to(dev)
pe_w = ot.annealing_weights(6, None)
rows = 64                                     # 64 image rows = 32 768 rays = 2.1 M samples per step (1/8 frame)
dirs = synth.pinhole_rays(H, W, device=dev, row0=H // 2 - rows // 2, rows=rows)
bi = torch.zeros(dirs.shape[0] * S, dtype=torch.long, device=dev)


GS, FM = build_ref.load("GridSamplerMine"), build_ref.load("FastMinv")
natives = "reference GridSamplerMine + FastMinv kernels (oracle/_ref)" if GS and FM else "torch F.grid_sample / torch inverse"


def lbs_inverse_ref(x):
    """ot.lbs_inverse with the reference's own native kernels where they are built."""
    if not (GS and FM):
        return ot.lbs_inverse(x, A, trans, ws, center, synth.BBOX_EXTEND, bi)
    nps = ((x - center.view(1, 3)) / synth.BBOX_EXTEND * 2).reshape(1, 1, 1, -1, 3).contiguous()
    w = GS.forward(ws, nps, 0, 1).view(24, -1).t()
    T = (w[:, :, None] * A[bi].reshape(-1, 24, 16)).sum(1).view(-1, 4, 4)
    Minv, ok = FM.Fast3x3Minv(T[:, :3, :3].contiguous())
    rhs = x - trans[bi] - T[:, :3, 3]
    return (Minv @ rhs[:, :, None])[:, :, 0], ok


def step():
    with torch.no_grad():
        x = (cam[None, None] + tk[None, :, None] * dirs[:, None, :]).reshape(-1, 3)
        xc, ok = lbs_inverse_ref(x)
        out = []
        for c in range(0, xc.shape[0], 65536):
            out.append(ot.sdf_mlp(xc[c:c + 65536], Ws, bs, pe_w)[0])
        return torch.cat(out)


for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    for _ in range(2):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 4
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"impl": "torch-gpu port of the reference path", "natives": natives, "matmul_tf32": tf32, "rays_per_s": dirs.shape[0] / (ms * 1e-3),
                      "ms_per_step": ms, "sample": f"{dirs.shape[0]} rays x {S} samples per step, 65 536-sample slabs"}))
