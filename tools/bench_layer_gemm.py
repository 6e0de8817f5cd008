"""One 131072 x 512 x 512 layer GEMM on planes in its epilogue variants, and the weight gradient of the same size."""
import sys, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, synth
dev = "cuda:0"
P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
g = synth.generator(1)
X = (torch.randn((P, 512), generator=g).abs() * 0.1).to(dev)
W = (torch.randn((512, 512), generator=g) * 0.06).to(dev)
b = torch.zeros((512,), device=dev)
xp = ops.split_planes(X, P, 512, 64.0)
wp = ops.split_planes(W, 512, 512, 1024.0)
Y = torch.empty((P, 512), device=dev)
yp = (torch.empty((P, 512), dtype=torch.float16, device=dev), torch.empty((P, 512), dtype=torch.float16, device=dev))
def timed(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("fwd none, fp32 out only      %.0f us" % timed(lambda: ops.mlp_layer_planes(xp, wp, P, 512, 512, 4, Y)))
print("fwd softplus + bias, fp32    %.0f us" % timed(lambda: ops.mlp_layer_planes(xp, wp, P, 512, 512, 5, Y, bias=b)))
print("fwd softplus + bias + planes %.0f us" % timed(lambda: ops.mlp_layer_planes(xp, wp, P, 512, 512, 5, Y, bias=b, y_planes=yp)))
dyn = torch.ones((1,), device=dev)
print("bwd softplus' + planes       %.0f us" % timed(lambda: ops.mlp_layer_planes(xp, wp, P, 512, 512, 1, Y, saved_input=X, dyn=dyn, a_has_dyn=True, y_planes=yp, planes_with_dyn=True)))
print("wgrad 512 x 512              %.0f us" % timed(lambda: ops.mlp_wgrad_planes(xp, xp, P, 512, 512, 1.0, dyn)))
print("N = 128 only (one column tile per row tile) fwd none %.0f us" % timed(lambda: ops.mlp_layer_planes(xp, wp, P, 128, 512, 4, Y)))
