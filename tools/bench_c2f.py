"""C3 of BASELINE.json: SDF network -> 257^3 grid through the coarse-to-fine sweep (33 -> 65 -> 129 -> 257) ->
marching cubes.  Times the whole extraction with the sweep as a device worklist (default when the query function is a
recmv_b200 network) and with the torch-op bookkeeping (any other query function), and marching cubes alone.
One JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import ops, testing  # noqa: E402
from recmv_b200.MCAcc import Seg3dLossless  # noqa: E402
from recmv_b200.discretize import discretize_sdf  # noqa: E402
from recmv_b200.model import getTmpSdf  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
for tag, pseed in (("geometric-init sphere", None), ("trained-like", 101)):
    sdf = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=pseed).to(dev)
    eng = Seg3dLossless(None, b_min=[-1, -1, -1], b_max=[1, 1, 1], resolutions=[33, 65, 129, 257], align_corners=False,
                        balance_value=0.0).to(dev)

    def run_device():
        return discretize_sdf(sdf, eng, None)

    def run_torch():
        def q(points):
            with torch.no_grad():
                return sdf.forward(points.reshape(-1, 3), None).reshape(1, 1, -1)
        eng.query_func = q
        eng.balance_value = 0.0
        with torch.no_grad():
            g = eng.forward()
        return ops.mc_gpu(g[0, 0].permute(2, 1, 0).contiguous(), eng.spacing_x, eng.spacing_y, eng.spacing_z, eng.bx,
                          eng.by, eng.bz, 0.0)
    res = {}
    for name, fn in (("device_worklist", run_device), ("torch_bookkeeping", run_torch)):
        for _ in range(3):
            v, f = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            v, f = fn()
        torch.cuda.synchronize()
        res[name] = {"ms_per_extraction": (time.perf_counter() - t0) / n * 1e3, "path": eng.last_sweep_path,
                     "queried_points": sum(s[3] for s in eng.stats), "levels": [(s[0], s[3]) for s in eng.stats],
                     "verts": int(v.shape[0]), "faces": int(f.shape[0])}
    g = eng.forward()
    vol = g[0, 0].permute(2, 1, 0).contiguous()
    for _ in range(3):
        ops.mc_gpu(vol, eng.spacing_x, eng.spacing_y, eng.spacing_z, eng.bx, eng.by, eng.bz, 0.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        ops.mc_gpu(vol, eng.spacing_x, eng.spacing_y, eng.spacing_z, eng.bx, eng.by, eng.bz, 0.0)
    torch.cuda.synchronize()
    res["marching_cubes_ms"] = (time.perf_counter() - t0) / 20 * 1e3
    out[tag] = res
print(json.dumps({"c3_extraction_257": out}))
