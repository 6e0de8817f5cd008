"""Stand-alone micro-benchmarks of the HBM-bound natives (K1/K2/K4 and the LBS kernels): achieved
ALGORITHMIC bytes / time against the measured copy bandwidth in MEASURED_PEAKS.json.  One JSON line per kernel.
Inputs are larger than L2 where the algorithm allows (the 181 MB skinning voxel is the gather target)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import ops, synth  # noqa: E402
from recmv_b200._lib import LAYOUT_NCDHW, LAYOUT_NDHWC  # noqa: E402

dev = torch.device("cuda", 0)
peak = 6592.2
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # L2 flush between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, alg_bytes, units, unit_name, note=""):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), unit_name + "_per_s": units / (ms * 1e-3),
                      "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                                   "frac": round(gbs / peak, 4)}, "algorithmic_bytes": alg_bytes, "note": note}))


g = synth.generator(3)
# ---- K1/K2: 3x3 inverse, 16.8 M matrices (72 B in + 37 B out) -------------------------------------------------
N = 1 << 24
m = torch.randn((N, 3, 3), generator=g).to(dev)
report("minv3x3_fwd f32", timed(lambda: ops.minv3x3(m)), N * (36 + 36 + 1), N, "matrices",
       "36 B in + 36 B inverse + 1 B flag per fp32 matrix (incl. output allocation)")
inv, ok = ops.minv3x3(m)
gr = torch.randn((N, 3, 3), generator=g).to(dev)
report("minv3x3_bwd f32", timed(lambda: ops.minv3x3_backward(gr, inv)), N * 108, N, "matrices", "2 x 36 B in + 36 B out")
del m, inv, gr
# ---- K4: grid sampler forward, the 24-channel 65x225x129 skinning voxel, 16.8 M points -------------------------
ws = synth.skinning_voxel((65, 225, 129), seed=7, device=dev)
ws_cl = ops.voxel_to_channels_last(ws)
P = 1 << 24
pts = ((torch.rand((P, 3), generator=g) - 0.5) * 2.0).to(dev)
grid = (pts / (synth.BBOX_EXTEND / 2)).view(1, 1, 1, P, 3)
alg = P * (8 * 24 * 4 + 12 + 96)
report("gridsample3d_fwd C=24 NCDHW (reference layout)", timed(lambda: ops.grid_sample3d_forward(ws, grid, LAYOUT_NCDHW), 5),
       alg, P, "points", "8 corners x 24 ch x 4 B gathered + 12 B coords + 96 B out per point")
report("gridsample3d_fwd C=24 NDHWC (channels-last copy)",
       timed(lambda: ops.grid_sample3d_forward(ws_cl.view(1, 65, 225, 129, 24), grid, LAYOUT_NDHWC), 5), alg, P, "points", "same")
# ---- A5 / A5': fused LBS forward and inverse, 16.8 M points -------------------------------------------------------
from recmv_b200.render import SdfRenderer  # noqa: E402
ren = SdfRenderer(dev, seed=0)
poses, trans = synth.poses_trans(1, seed=11)
A, t = ren.bone_matrices(poses.to(dev), trans.to(dev))
alg = P * (8 * 24 * 4 + 12 + 12)
report("lbs_fwd (sample + blend + apply)", timed(lambda: ops.lbs_forward(pts, A, t, ren.ws_cl, synth.BBOX_CENTER, synth.BBOX_EXTEND, None, P), 5),
       alg, P, "points", "768 B gathered + 12 B in + 12 B out per point")
report("lbs_inverse (sample + blend + 3x3 inverse + apply)",
       timed(lambda: ops.lbs_inverse(pts, A, t, ren.ws_cl, synth.BBOX_CENTER, synth.BBOX_EXTEND, None, P), 5),
       alg + P, P, "points", "768 B gathered + 12 B in + 13 B out per point")
