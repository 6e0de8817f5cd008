import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recmv_b200.model as M
from recmv_b200 import synth as sy, ops, _lib
from recmv_b200.render import SdfRenderer
dev = torch.device("cuda", 0)
P = int(sys.argv[1]); N = int(sys.argv[2]); which = sys.argv[3] if len(sys.argv) > 3 else "jac"
ren = SdfRenderer(dev, seed=0)
g = sy.generator(21)
pts = ((torch.rand((P, 3), generator=g) - 0.5) * 1.2).to(dev)
conds = (torch.randn((1, 128), generator=g) * 0.1).to(dev)
poses, trans = sy.poses_trans(1, seed=11)
poses, trans = poses.to(dev), trans.to(dev)
torch.manual_seed(3)
tr = M.MLPTranslator(128, 6).to(dev)
deformer = M.CompositeDeformer([tr, ren.skinner])
ratio = {"sdfRatio": None, "deformerRatio": None, "renderRatio": None}
bi = torch.zeros((P,), dtype=torch.long, device=dev)
ref = None
with torch.no_grad():
    for i in range(N):
        try:
            if which == "trjac":   # translator-only forward-mode launch (no LBS stage)
                res = ops.deformer_forward(pts, conds, tr.packed_weights(), [1.0] * 12, bi, 0, None, None, True, True, True)
                d, J = res[0], res[3]
            elif which == "sdfjac":
                d, J = ren.sdf_net.value_and_grad(pts, None)
            elif which == "jac":
                d, J = deformer.value_and_jacobian(pts, [conds, [poses, trans]], bi, ratio=ratio, offset_type="body")
            else:
                d = deformer(pts, [conds, [poses, trans]], bi, ratio=ratio, offset_type="body"); J = d
            if i % 8 == 7 or i == N - 1:
                torch.cuda.synchronize()
                chk = (float(d.double().sum()), float(J.double().sum()))
                if ref is None: ref = chk
                if chk != ref: print("MISMATCH at", i, chk, ref)
        except Exception as e:
            info = (__import__("ctypes").c_int * 3)()
            print("FAILED at launch", i, repr(e)[:120], "status", _lib.load().recmv_check_async_errors(info, 0), list(info))
            sys.exit(1)
print("ok", N, "launches", ref)
