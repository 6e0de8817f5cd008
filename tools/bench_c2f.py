"""C3 of BASELINE.json: SDF network -> 257^3 grid through the coarse-to-fine sweep (33 -> 65 -> 129 -> 257) ->
marching cubes; time split between network queries, sweep bookkeeping and MC."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import ops  # noqa: E402
from recmv_b200.MCAcc import Seg3dLossless  # noqa: E402
from recmv_b200.discretize import discretize_sdf  # noqa: E402
from recmv_b200.model import getTmpSdf  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
sdf = getTmpSdf(dev, 6, 0.6, 256)
eng = Seg3dLossless(None, b_min=[-1, -1, -1], b_max=[1, 1, 1], resolutions=[33, 65, 129, 257], align_corners=False,
                    balance_value=0.0).to(dev)
qtime = [0.0]
orig_forward = sdf.forward


def timed_forward(x, ratio):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = orig_forward(x, ratio)
    torch.cuda.synchronize()
    qtime[0] += time.perf_counter() - t0
    return out


for rep in range(3):
    sdf.forward = orig_forward
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, f = discretize_sdf(sdf, eng, None)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    sdf.forward = timed_forward
    qtime[0] = 0.0
    v, f = discretize_sdf(sdf, eng, None)
    torch.cuda.synchronize()
    queried = sum(s[3] for s in eng.stats)
    print(f"rep {rep}: total {total * 1e3:.1f} ms (untimed-inside) | network queries {qtime[0] * 1e3:.1f} ms for {queried} points "
          f"({queried / 257 ** 3 * 100:.1f} % of the dense grid), levels {[(s[0], s[3]) for s in eng.stats]} | V={v.shape[0]} F={f.shape[0]}")
g = eng.forward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
vol = g[0, 0].permute(2, 1, 0).contiguous()
e0.record()
for _ in range(10):
    ops.mc_gpu(vol, eng.spacing_x, eng.spacing_y, eng.spacing_z, eng.bx, eng.by, eng.bz, 0.0)
e1.record()
torch.cuda.synchronize()
print(f"marching cubes alone: {e0.elapsed_time(e1) / 10:.3f} ms per call")
