import sys, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, synth, testing
import recmv_b200.model as M
from torch.profiler import profile, ProfilerActivity
dev = "cuda:0"
P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
net = testing.build_sdf(M.getTmpSdf, seed=0, perturb_seed=101).to(dev)
g = synth.generator(5)
x = ((torch.rand((P, 3), generator=g) - 0.5) * 1.2).to(dev)
c0 = (torch.randn((P, 1), generator=g) / P).to(dev); c1 = (torch.randn((P, 256), generator=g) / P * 0.1).to(dev)
def step():
    net.zero_grad(set_to_none=True)
    xg = x.detach().requires_grad_(True)
    y = net(xg, None)
    ((y * c0).sum() + (net.rendcond * c1).sum()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=64))
