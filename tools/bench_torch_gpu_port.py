"""Informative only (profiles/): the reference-style PyTorch path -- separate torch ops, fp32 cuBLAS GEMMs, every
activation through HBM -- for the same workload ON THE SAME B200, i.e. the oracle's restatement moved to cuda:0
("the reference path on the same B200 ... the reference single-GPU rays/s the >= 10x target is measured
against").  The reference's own CUDA extensions cannot be built for sm_100 without patching, so torch's
F.grid_sample stands in for its sampler here (what its own check script compares against)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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


def step():
    with torch.no_grad():
        x = (cam[None, None] + tk[None, :, None] * dirs[:, None, :]).reshape(-1, 3)
        xc, ok = ot.lbs_inverse(x, A, trans, ws, center, synth.BBOX_EXTEND, bi)
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
    print(json.dumps({"impl": "torch-gpu port of the reference path", "matmul_tf32": tf32, "rays_per_s": dirs.shape[0] / (ms * 1e-3),
                      "ms_per_step": ms, "sample": f"{dirs.shape[0]} rays x {S} samples per step, 65 536-sample slabs"}))
