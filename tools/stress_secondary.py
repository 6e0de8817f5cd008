"""Stress the secondary fused launches back to back (as bench.py's secondary block does) and report the device
status record after every step."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from recmv_b200 import ops, _lib
from recmv_b200.render import SdfRenderer
dev = torch.device("cuda", 0)
ren = SdfRenderer(dev, mode=_lib.MLP_TC_F16X3, samples=64)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    try:
        out = bench.secondary_rates(dev, ren, _lib.MLP_TC_F16X3)
        torch.cuda.synchronize()
        ops.check_async_errors()
        print(rep, {k.split(" ")[0]: round(v["ms_per_1M_points"], 2) for k, v in out.items()})
    except Exception as e:
        print("FAILED rep", rep, repr(e)[:300])
        info = (__import__("ctypes").c_int * 3)()
        print("status", _lib.load().recmv_check_async_errors(info, 0), list(info))
        break
