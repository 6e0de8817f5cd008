import sys, time, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, testing
from recmv_b200.MCAcc import Seg3dLossless
from recmv_b200.discretize import discretize_sdf
from recmv_b200.model import getTmpSdf
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
sdf = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=None).to(dev)
eng = Seg3dLossless(None, b_min=[-1, -1, -1], b_max=[1, 1, 1], resolutions=[33, 65, 129, 257], align_corners=False, balance_value=0.0).to(dev)
for _ in range(3): discretize_sdf(sdf, eng, None)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): discretize_sdf(sdf, eng, None)
torch.cuda.synchronize()
print("ms per extraction", (time.perf_counter() - t0) / 10 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): discretize_sdf(sdf, eng, None)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
