import sys, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, synth, testing
import recmv_b200.model as M
DEV = "cuda:0"
net = testing.build_sdf(M.getTmpSdf, seed=0, perturb_seed=101).to(DEV)
def merr(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))
for P in (1024, 1500, 777):
    x0 = ((torch.rand((P, 3), generator=synth.generator(3)) - 0.5) * 1.2).to(DEV)
    c = (torch.randn((P, 256), generator=synth.generator(4)) / P).to(DEV)
    w = torch.randn((P, 3), generator=synth.generator(5)).to(DEV)
    res = {}
    for fused in (False, True):
        net.train_fused = fused
        for case in ("feat only", "feat after create_graph", "feat + normals"):
            net.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_(True)
            y = net(x, {'sdfRatio': 0.8})
            feat = net.rendcond
            loss = (feat * c).sum()
            if case != "feat only":
                with ops.input_grad_only():
                    (gx,) = torch.autograd.grad(y, x, torch.ones_like(y), retain_graph=True, create_graph=True)
                if case == "feat + normals":
                    loss = loss + (torch.nn.functional.normalize(gx, dim=1) * w).sum() / P
            loss.backward()
            res[(fused, case)] = (x.grad.clone(), net.lin3.weight_v.grad.clone())
    for case in ("feat only", "feat after create_graph", "feat + normals"):
        print(P, case, "dx", f"{merr(res[(True, case)][0], res[(False, case)][0]):.1e}", "dW3", f"{merr(res[(True, case)][1], res[(False, case)][1]):.1e}")
