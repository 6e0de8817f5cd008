"""tcgen05 issue-rate microbenchmark: cycles per MMA for candidate tile shapes, issue-thread overheads and
background loads (flag bits: see csrc/tc_microbench.cu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import _lib

lib = _lib.load_diag()
dev = torch.device("cuda", 0)
out = torch.zeros(8, dtype=torch.int64, device=dev)
gsrc = torch.randint(0, 255, (148 * 65536 + 65536,), dtype=torch.uint8, device=dev)
iters = 512
NAMES = {1: "alt-acc", 2: "epi-warps", 4: "bulk-copies", 8: "mcast-commit", 16: "commit/4", 32: "wait-blocking",
         64: "wait-pipelined", 128: "two-issuers", 256: "commit/8", 512: "commit/12"}


def run(cg, M, N, ctas, flags=0):
    for rep in range(2):
        out.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.recmv_tc_microbench(cg, M, N, iters, ctas, flags, gsrc.data_ptr(), out.data_ptr(), None)
        e1.record()
        torch.cuda.synchronize()
    assert rc == 0, (rc, flags)
    o = out.cpu().tolist()
    n = iters * 32
    ms = e0.elapsed_time(e1)
    tf = 2.0 * M * N * 16 * n * (ctas // cg) / (ms * 1e-3) / 1e12
    cyc = max(o[0], o[4]) / n
    desc = "+".join(v for k, v in NAMES.items() if flags & k) or "plain"
    print(f"{cg}  {M:4d} {N:4d} {ctas:4d}  {cyc:8.1f}  {ms:7.3f}  {tf:8.1f}   {desc}")


print("cg  M    N   ctas  cyc/MMA      ms    TFLOP/s(chip)  variant")
if len(sys.argv) > 1 and sys.argv[1] == "shapes":
    for cg, M, N in [(2, 128, 256), (2, 256, 128), (2, 256, 256), (2, 128, 128), (2, 256, 64), (1, 128, 256),
                     (1, 64, 256), (1, 128, 128)]:
        for ctas in (cg, 148):
            run(cg, M, N, ctas)
else:
    for fl in (0, 1, 2, 4, 16, 24, 32, 64, 256, 512, 16 + 32, 16 + 64, 256 + 64, 512 + 64, 128, 128 + 16, 128 + 16 + 32,
               128 + 16 + 64, 128 + 256 + 64, 2 + 16, 2 + 16 + 32, 2 + 16 + 64, 2 + 256 + 64, 2 + 512 + 64, 2 + 128,
               2 + 128 + 16 + 32, 2 + 128 + 16 + 64, 2 + 128 + 256 + 64, 2 + 4 + 128 + 16 + 64 + 8, 2 + 4 + 512 + 64 + 8):
        run(2, 128, 256, 148, fl)
