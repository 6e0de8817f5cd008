import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recmv_b200.model as M
from recmv_b200 import synth as sy, ops
from recmv_b200.render import SdfRenderer
dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 32 * 74 * 3
small = len(sys.argv) > 2
ren = SdfRenderer(dev, seed=0, voxel_shape=(17, 33, 21) if small else (65, 225, 129))
g = sy.generator(21)
pts = ((torch.rand((P, 3), generator=g) - 0.5) * 1.2).to(dev)
conds = (torch.randn((1, 128), generator=g) * 0.1).to(dev)
poses, trans = sy.poses_trans(1, seed=11)
torch.manual_seed(3)
tr = M.MLPTranslator(128, 6).to(dev)
deformer = M.CompositeDeformer([tr, ren.skinner])
ratio = {"sdfRatio": None, "deformerRatio": None, "renderRatio": None}
bi = torch.zeros((P,), dtype=torch.long, device=dev)
with torch.no_grad():
    d0 = deformer(pts, [conds, [poses.to(dev), trans.to(dev)]], bi, ratio=ratio, offset_type="body")
    torch.cuda.synchronize()
    print("fwd ok")
    d1, J = deformer.value_and_jacobian(pts, [conds, [poses.to(dev), trans.to(dev)]], bi, ratio=ratio, offset_type="body")
    torch.cuda.synchronize()
print("jac ok", (d0 - d1).abs().max().item(), J.abs().max().item())
