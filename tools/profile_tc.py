"""One warm + N measured launches of the fused render kernel on a fraction of the 512x512x64 frame
(for `ncu`): python tools/profile_tc.py [tc3|tc1|simt] [rows] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import synth  # noqa: E402
from recmv_b200.render import SdfRenderer  # noqa: E402

mode = {"simt": 0, "tc3": 1, "tc1": 2}[sys.argv[1] if len(sys.argv) > 1 else "tc3"]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
ren = SdfRenderer(dev, mode=mode, samples=64)
poses, trans = synth.poses_trans(1, seed=11)
A, t = ren.bone_matrices(poses.to(dev), trans.to(dev))
dirs = synth.pinhole_rays(512, 512, device=dev, row0=256 - rows // 2, rows=rows)
for _ in range(1 + reps):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    ren.render(dirs, A, t)
    ev1.record()
    torch.cuda.synchronize()
    print(f"rows={rows} mode={mode}: {ev0.elapsed_time(ev1):.2f} ms -> {dirs.shape[0] / ev0.elapsed_time(ev1) * 1e3:.0f} rays/s")
