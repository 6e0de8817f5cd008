"""Small workloads for ncu captures (round 2): `python tools/prof_r2.py mc|train|render|c2f`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import ops, synth, testing  # noqa: E402
from recmv_b200.model import getTmpSdf  # noqa: E402

dev = torch.device("cuda", 0)
what = sys.argv[1]
if what == "mc":
    grid = synth.sphere_sdf_grid(257, num=8, seed=3, device=dev)
    for _ in range(3):
        ops.mc_gpu(grid, 2 / 256, 2 / 256, 2 / 256, -1.0, -1.0, -1.0, 0.0)
elif what == "train":
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    net = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=101).to(dev)
    x = ((torch.rand((P, 3), generator=synth.generator(1)) - 0.5) * 1.2).to(dev)
    for _ in range(2):
        xg = x.clone().requires_grad_(True)
        y = net(xg, None)
        ((y.sum() + net.rendcond.sum() * 0.1) / P).backward()
elif what == "render":
    from recmv_b200.render import SdfRenderer
    ren = SdfRenderer(dev, samples=64)
    poses, trans = synth.poses_trans(1, seed=11)
    A, t = ren.bone_matrices(poses.to(dev), trans.to(dev))
    dirs = synth.pinhole_rays(512, 512, device=dev)
    for _ in range(2):
        ren.render(dirs, A, t)
elif what == "c2f":
    from recmv_b200.MCAcc import Seg3dLossless
    from recmv_b200.discretize import discretize_sdf
    sdf = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=101).to(dev)
    eng = Seg3dLossless(None, b_min=[-1, -1, -1], b_max=[1, 1, 1], resolutions=[33, 65, 129, 257], align_corners=False,
                        balance_value=0.0).to(dev)
    for _ in range(2):
        discretize_sdf(sdf, eng, None)
torch.cuda.synchronize()
ops.check_async_errors()
