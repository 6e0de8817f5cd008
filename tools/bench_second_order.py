"""Second-order terms of the training step: fused (tcgen05 GEMMs) vs torch autograd over cuBLAS fp32, same GPU.
Usage: python tools/bench_second_order.py [points]"""
import json, sys, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, synth, testing, utils
import recmv_b200.model as M
dev = "cuda:0"
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.backends.cuda.matmul.allow_tf32 = False
net = testing.build_sdf(M.getTmpSdf, seed=0, perturb_seed=101).to(dev)
torch.manual_seed(1)
tr = testing.perturb_module(M.MLPTranslator(128, 6), 202, scale=0.5).to(dev)
g = synth.generator(5)
x = ((torch.rand((P, 3), generator=g) - 0.5) * 1.2).to(dev)
conds = (torch.randn((2, 128), generator=g) * 0.1).to(dev).requires_grad_(True)
p2 = x[: (P // 2) * 2].view(2, -1, 3)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def eik():
    net.zero_grad(set_to_none=True)
    utils.eikonal_loss(net, x.clone(), None).backward()


def regu():
    tr.zero_grad(set_to_none=True)
    utils.deformation_regulariser(tr, p2.clone(), conds, {"deformerRatio": 0.6}, 0.2, offset_type="body").backward()


def regu_host_svd():      # the reference's formulation: singular values on the host
    tr.zero_grad(set_to_none=True)
    pts = p2.clone().requires_grad_()
    J = utils.compute_Jacobian(pts, tr(pts, conds, ratio={"deformerRatio": 0.6}, offset_type="body"), True, True)
    _, s, _ = torch.svd(J.cpu())
    s = torch.log(s.to(dev))
    utils.GMRobustError((s * s).sum(1), 0.2, True).mean().backward()


out = {"points": P}
out["eikonal_fused_ms"] = timed(eik)
out["regulariser_fused_ms"] = timed(regu)
net.train_fused = False
tr.train_fused = False
out["eikonal_torch_ms"] = timed(eik)
out["regulariser_torch_device_svd_ms"] = timed(regu)
out["regulariser_torch_host_svd_ms"] = timed(regu_host_svd)
print(json.dumps(out))
