"""Per-layer signed error of the training forward given OUR layer inputs (isolates each GEMM + epilogue)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from recmv_b200 import ops, synth, testing
from recmv_b200.model import getTmpSdf
dev = "cuda:0"
def proj(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (b @ b) - 1.0)
net = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=101).to(dev)
P = 4096
x = (torch.rand((P, 3), generator=synth.generator(3)) * 1.4 - 0.7).to(dev)
Ws, bs = [], []
for l in range(9):
    lin = getattr(net, f"lin{l}")
    Ws.append((lin.weight_v * (lin.weight_g / lin.weight_v.norm(dim=1, keepdim=True))).detach().contiguous())
    bs.append(lin.bias.detach().contiguous())
act = [torch.zeros((P, 64), device=dev)] + [torch.zeros((P, 512), device=dev) for _ in range(8)]
ops.pe_forward(x, [1.0] * 12, 6, act[0], act[4][:, 473:])
pe64 = ops._pe_torch(x.double(), [1.0] * 12, 6)
print("PE proj", f"{proj(act[0][:, :39], pe64):+.2e}", "max", float((act[0][:, :39].double() - pe64).abs().max()))
for l in range(8):
    o, i = Ws[l].shape
    ops.mlp_fwd_layer(act[l], Ws[l], bs[l], o, i, ops.ACT_SOFTPLUS100, act[l + 1], pre_scale=0.7071067811865476 if l == 4 else 1.0)
    xin = act[l][:, :i].double() * (0.7071067811865476 if l == 4 else 1.0)
    mm = xin @ Ws[l].double().T
    z = mm + bs[l].double()
    a = torch.nn.functional.softplus(z, beta=100)
    Z = torch.empty((P, 512), device=dev)
    ops.mlp_fwd_layer(act[l], Ws[l], None, o, i, ops.ACT_NONE, Z, pre_scale=0.7071067811865476 if l == 4 else 1.0)
    Zb = torch.empty((P, 512), device=dev)
    ops.mlp_fwd_layer(act[l], Ws[l], bs[l], o, i, ops.ACT_NONE, Zb, pre_scale=0.7071067811865476 if l == 4 else 1.0)
    dz = (Zb[:, :o].double() - z)
    print(f"layer {l}: matmul proj {proj(Z[:, :o], mm):+.2e}  z proj {proj(Zb[:, :o], z):+.2e}  mean dz {float(dz.mean()):+.2e} rms dz {float(dz.pow(2).mean().sqrt()):.2e} "
          f"rms z {float(z.pow(2).mean().sqrt()):.2e}  softplus proj {proj(act[l + 1][:, :o], a):+.2e}  "
          f"a32-of-z64 proj {proj(torch.nn.functional.softplus(z.float(), beta=100), a):+.2e}")
# growth of the signed error along the chain (ours vs the float64 chain from the same x)
h = pe64
pe = pe64
for l in range(9):
    if l == 4:
        h = torch.cat([h, pe], 1) * 0.7071067811865476
    h = torch.nn.functional.linear(h, Ws[l].double(), bs[l].double())
    if l < 8:
        h = torch.nn.functional.softplus(h, beta=100)
        o = Ws[l].shape[0]
        d = act[l + 1][:, :o].double() - h
        big = h > 1e-3
        print(f"chain after layer {l}: proj {proj(act[l + 1][:, :o], h):+.2e}  mean rel err on a > 1e-3: {float((d[big] / h[big]).mean()):+.2e}  rms rel {float((d[big] / h[big]).pow(2).mean().sqrt()):.2e}")
sdf = torch.empty((P, 1), device=dev); feat = torch.empty((P, 256), device=dev)
ops.mlp_fwd_layer(act[8], Ws[8], bs[8], 257, 512, ops.ACT_NONE, sdf, split=1, Y2=feat)
print(f"chain output: sdf proj {proj(sdf, h[:, :1]):+.2e} feat proj {proj(feat, h[:, 1:]):+.2e}; mean abs err sdf {float((sdf.double() - h[:, :1]).mean()):+.2e} rms sdf {float(h[:, :1].pow(2).mean().sqrt()):.2e}")
mm8 = act[8].double() @ Ws[8].double().T
print("last layer: |matmul part| rms", float(mm8[:, 0].pow(2).mean().sqrt()), "bias", float(bs[8][0]))
