import sys, time, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, testing
from recmv_b200.MCAcc import Seg3dLossless
from recmv_b200.discretize import discretize_sdf
from recmv_b200.model import getTmpSdf
dev = torch.device("cuda", 0)
for pseed in (None, 101, None, 101):
    sdf = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=pseed).to(dev)
    eng = Seg3dLossless(None, b_min=[-1, -1, -1], b_max=[1, 1, 1], resolutions=[33, 65, 129, 257], align_corners=False, balance_value=0.0).to(dev)
    ts = []
    for rep in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        v, f = discretize_sdf(sdf, eng, None)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(pseed, eng.last_sweep_path, [round(t, 2) for t in ts], v.shape[0], eng.stats)
