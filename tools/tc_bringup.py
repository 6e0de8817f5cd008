"""Bring-up diagnostics for the tcgen05 SDF kernel (run on the GPU box):
per-layer raw accumulators of the first tile against fp32/fp64 torch references, the kernel's
bounded-wait status record, and end-to-end errors for both precision modes.  Writes a text report and
an .npz of mismatching blocks into gpurun_out/.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_torch as ot  # noqa: E402
from recmv_b200 import _lib, ops, synth, testing  # noqa: E402
from recmv_b200.model import getTmpSdf  # noqa: E402


def run_debug(x, packed, P, passes, layer):
    lib = _lib.load()
    dev = x.device
    sdf = torch.full((P,), float("nan"), device=dev)
    feat = torch.full((P, 256), float("nan"), device=dev)
    dbg = torch.full((128, 512), float("nan"), device=dev)
    st = (ctypes.c_int * 4)()
    pe = (ctypes.c_float * 12)(*([1.0] * 12))
    rc = lib.recmv_sdf_mlp_tc_debug(x.data_ptr(), packed.data_ptr(), pe, sdf.data_ptr(), feat.data_ptr(), P,
                                    passes, layer, dbg.data_ptr(), st, None, None)
    torch.cuda.synchronize()
    return rc, list(st), sdf, feat, dbg


def main():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    rep = open(os.path.join(out_dir, "tc_bringup.txt"), "w")

    def say(*a):
        msg = " ".join(str(t) for t in a)
        print(msg)
        rep.write(msg + "\n")
        rep.flush()

    dev = torch.device("cuda", 0)
    net = testing.build_sdf(getTmpSdf, seed=0, perturb_seed=101).to(dev)
    packed = net.packed_weights()
    Ws, bs = net.effective_weights()
    Ws = [w.detach().double().cpu() for w in Ws]
    bs = [b.detach().double().cpu() for b in bs]
    Ws[4] = Ws[4] / np.sqrt(2)
    g = synth.generator(1234)
    x = (torch.rand((4096, 3), generator=g) * 1.2 - 0.6)
    xd = x.to(dev)
    # fp64 reference of every layer's pre-activation for the first 128 points
    inp = ot.embed(x[:128].double(), 6, None)
    pre, h = [], inp
    for l in range(9):
        if l == 4:
            h = torch.cat([h, inp], 1)
        z = h @ Ws[l].t() + bs[l]
        pre.append(z)
        h = torch.nn.functional.softplus(z, beta=100) if l < 8 else z
    dumps = {}
    for passes in (1, 3):
        for layer in (0, 1, 4, 8):
            rc, st, sdf, feat, dbg = run_debug(xd, packed, 128, passes, layer)
            raw = (pre[layer] - bs[layer])  # accumulator excludes the bias
            n = raw.shape[1]
            nc = min(n, 512) if layer < 8 else 256
            got = dbg[:, :nc].double().cpu()
            ref = raw[:, :nc]
            err = (got - ref).abs()
            finite = torch.isfinite(got)
            say(f"passes={passes} layer={layer} rc={rc} status={st} finite={finite.float().mean():.3f} "
                f"max_abs_err={err[finite].max().item() if finite.any() else float('nan'):.3e} "
                f"ref_absmax={ref.abs().max():.3e}")
            if not finite.all() or err[finite].max() > 1e-2 * ref.abs().max():
                # locate the structure of the mismatch: per 64-row x 64-col block
                nn_ = (min(n, 512) // 64) * 64
                blk = err[:, :nn_].nan_to_num(1e9).reshape(2, 64, -1, 64).amax(dim=(1, 3))
                say("  block max err (rows 0-63 / 64-127 x 64-col blocks):", [[f"{v:.1e}" for v in r] for r in blk.tolist()])
                dumps[f"p{passes}_l{layer}_got"] = got.numpy()
                dumps[f"p{passes}_l{layer}_ref"] = ref.numpy()
        rc, st, sdf, feat, _ = run_debug(xd, packed, 4096, passes, -1)
        s_ref = pre[8][:, 0]
        e = ((sdf[:128].double().cpu() - s_ref).abs() / s_ref.abs().clamp_min(1e-2)).max().item()
        with torch.no_grad():
            net.mlp_mode = _lib.MLP_FP32_SIMT
            y = net(xd, None)
        eall = ((sdf - y[:, 0]).abs() / y[:, 0].abs().clamp_min(1e-2)).max().item()
        fe = ((feat - net.rendcond).abs() / net.rendcond.abs().clamp_min(1e-2)).max().item()
        say(f"passes={passes} end-to-end P=4096 rc={rc} status={st} rel_err(sdf, fp64 ref, first 128)={e:.3e} "
            f"rel_err(sdf vs simt, all)={eall:.3e} rel_err(feat vs simt)={fe:.3e}")
    if dumps:
        np.savez_compressed(os.path.join(out_dir, "tc_bringup_dumps.npz"), **dumps)
    rep.close()


if __name__ == "__main__":
    main()
