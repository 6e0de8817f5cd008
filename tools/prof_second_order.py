import sys, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, synth, testing, utils
import recmv_b200.model as M
from torch.profiler import profile, ProfilerActivity
dev = "cuda:0"
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
net = testing.build_sdf(M.getTmpSdf, seed=0, perturb_seed=101).to(dev)
x = ((torch.rand((P, 3), generator=synth.generator(5)) - 0.5) * 1.2).to(dev)
def eik():
    net.zero_grad(set_to_none=True)
    utils.eikonal_loss(net, x.clone(), None).backward()
for _ in range(3): eik()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): eik()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
