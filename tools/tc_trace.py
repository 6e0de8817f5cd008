"""Cycle-stamp timeline of the tcgen05 kernel (cluster 0, leader CTA, first two tiles)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import _lib, synth, testing
from recmv_b200.model import getTmpSdf

dev = torch.device("cuda", 0)
net = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=101).to(dev)
packed = net.packed_weights()
P = 128 * 74 * 4
x = (torch.rand((P, 3), generator=synth.generator(1)) * 1.2 - 0.6).to(dev)
lib = _lib.load()
for passes in (3, 1):
    sdf = torch.empty(P, device=dev)
    trace = torch.zeros(4 * 2 * 9 * 16, dtype=torch.int64, device=dev)
    st = (ctypes.c_int * 4)()
    pe = (ctypes.c_float * 12)(*([1.0] * 12))
    for rep in range(2):
        trace.zero_()
        rc = lib.recmv_sdf_mlp_tc_debug(x.data_ptr(), packed.data_ptr(), pe, sdf.data_ptr(), None, P, passes, -1, None,
                                        st, trace.data_ptr(), None)
        torch.cuda.synchronize()
    t = trace.cpu().view(4, 2, 9, 16)
    t0 = int(t[0, 0, 0, 0])
    print(f"=== passes={passes} rc={rc} status={list(st)}  (cycles relative to MMA start of tile 0)")
    for it in range(2):
        for l in range(9):
            m = [int(v) - t0 if v else -1 for v in t[0, it, l, :6]]
            e1 = [int(v) - t0 if v else -1 for v in t[1, it, l, :8]]
            e2 = [int(v) - t0 if v else -1 for v in t[2, it, l, :5]]
            pr = [int(v) - t0 if v else -1 for v in t[3, it, l, :2]]
            ws = [int(v) for v in t[0, it, l, 6:12]]
            if any(ws):
                print(f"      issuer0: slot-wait cycles={ws[0]} blocking={ws[1]} kblock-wait={ws[2]} | issuer1: {ws[3]} {ws[4]} {ws[5]}")
            print(f"it{it} L{l} MMA[start,accfree,kb0rdy,kb4rdy,lastkb,commit]={m}  EPIw4[wait,go,nt0,nt1,done,c0:ld,c0:stored,c0:arrived]={e1}  "
                  f"EPIw11={e2} PROD[start,lastkb]={pr}")
