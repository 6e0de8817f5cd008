#!/usr/bin/env python
"""TEST INFRASTRUCTURE (runs on the B200 box): executes the reference's OWN native extensions -- built
unmodified for sm_100a by oracle/build_ref.py into oracle/_ref/ -- and

  1. writes golden vectors of the reference marching cubes (canonical vertex / face order, because the
     reference emits through atomics) for the grids the tests use, incl. the three anisotropic production
     pyramids of train.py:47-71 and 257^3            -> gpurun_out/ref_golden/mc_ref.npz
  2. compares this package's kernels with them in the same process (printed + JSON)
  3. times reference natives and ours side by side (mc_gpu, Fast3x3Minv, GridSamplerMine.forward,
     interp2x_boundary3d.forward)                     -> gpurun_out/ref_natives.json

The committed copies live in tests/golden/mc_ref.npz and profiles/r02_ref_natives.json.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, mc_oracle  # noqa: E402
from recmv_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(os.path.join(OUT, "ref_golden"), exist_ok=True)

MC_CASES = [  # name, shape, spheres, seed, iso, keep the full arrays?
    ("s41", (41, 41, 41), 4, 3, 0.0, True),
    ("a21x37x13", (21, 37, 13), 4, 3, 0.0, True),
    ("a33x17x50", (33, 17, 50), 4, 3, 0.0, True),
    ("s41_iso0.1", (41, 41, 41), 4, 3, 0.1, True),
    ("coarse225x321x129", (225, 321, 129), 8, 5, 0.0, False),
    ("medium289x385x193", (289, 385, 193), 8, 6, 0.0, False),
    ("fine321x417x225", (321, 417, 225), 8, 7, 0.0, False),
    ("s257", (257, 257, 257), 8, 3, 0.0, False),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    report = {"gpu": torch.cuda.get_device_name(0)}
    golden = {}
    MCGpu = build_ref.load("MCGpu")
    if MCGpu is None:
        raise SystemExit("oracle/_ref/MCGpu*.so missing: run oracle/build_ref.py in the container first")
    MCGpu.mc_init(0)
    mc_rows = []
    for name, shape, num, seed, iso, keep in MC_CASES:
        sdf = synth.sphere_sdf_grid(shape, num=num, seed=seed, device=dev)
        step = tuple(2.0 / (n - 1) for n in shape)
        rv, rf = MCGpu.mc_gpu(sdf, *step, -1.0, -1.0, -1.0, iso)
        ov, of = ops.mc_gpu(sdf, *step, -1.0, -1.0, -1.0, iso)
        torch.cuda.synchronize()
        crv, crf = mc_oracle.canonical(rv.cpu().numpy(), rf.cpu().numpy())
        cov, cof = mc_oracle.canonical(ov.cpu().numpy(), of.cpu().numpy())
        same_counts = crv.shape == cov.shape and crf.shape == cof.shape
        faces_equal = bool(same_counts and np.array_equal(crf, cof))
        verts_bits = bool(same_counts and np.array_equal(crv, cov))
        vmax = float(np.abs(crv - cov).max()) if same_counts and len(crv) else None
        row = {"case": name, "shape": list(shape), "iso": iso, "V": int(crv.shape[0]), "F": int(crf.shape[0]),
               "ours_V": int(cov.shape[0]), "ours_F": int(cof.shape[0]), "faces_bit_exact": faces_equal,
               "verts_bit_identical": verts_bits, "verts_max_abs_diff": vmax}
        mc_rows.append(row)
        print("MC", row, flush=True)
        golden[name + "_shape"] = np.asarray(shape, np.int64)
        golden[name + "_meta"] = np.asarray([num, seed], np.int64)
        golden[name + "_iso"] = np.asarray([iso], np.float32)
        golden[name + "_counts"] = np.asarray([crv.shape[0], crf.shape[0]], np.int64)
        golden[name + "_faces_sha256"] = np.frombuffer(bytes.fromhex(sha(crf.astype(np.int64))), dtype=np.uint8)
        golden[name + "_verts_sum"] = crv.astype(np.float64).sum(0)
        golden[name + "_verts_abs_sum"] = np.abs(crv.astype(np.float64)).sum(0)
        if keep:
            golden[name + "_verts"] = crv.astype(np.float32)
            golden[name + "_faces"] = crf.astype(np.int64)
        else:   # a deterministic 4096-row sample of the canonical arrays
            iv = np.linspace(0, crv.shape[0] - 1, 4096).astype(np.int64)
            jf = np.linspace(0, crf.shape[0] - 1, 4096).astype(np.int64)
            golden[name + "_verts_idx"] = iv
            golden[name + "_verts_sample"] = crv[iv].astype(np.float32)
            golden[name + "_faces_idx"] = jf
            golden[name + "_faces_sample"] = crf[jf].astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "ref_golden", "mc_ref.npz"), **golden)
    report["mc_parity"] = mc_rows

    # ---- timings: reference natives vs ours, same inputs, same process -------------------------------------------
    t = {}
    sdf = synth.sphere_sdf_grid(257, num=8, seed=3, device=dev)
    st = 2 / 256
    t["mc_gpu 257^3 (ms/call)"] = {
        "reference": timed(lambda: MCGpu.mc_gpu(sdf, st, st, st, -1.0, -1.0, -1.0, 0.0)),
        "ours": timed(lambda: ops.mc_gpu(sdf, st, st, st, -1.0, -1.0, -1.0, 0.0))}
    del sdf
    FastMinv = build_ref.load("FastMinv")
    g = synth.generator(3)
    for n in (262144, 1 << 24):
        m = torch.randn((n, 3, 3), generator=g).to(dev)
        ri, rc = FastMinv.Fast3x3Minv(m)
        oi, oc = ops.minv3x3(m)
        both = rc.bool() & oc
        t[f"Fast3x3Minv n={n} (ms/call)"] = {
            "reference": timed(lambda: FastMinv.Fast3x3Minv(m)), "ours": timed(lambda: ops.minv3x3(m)),
            "flags_equal": bool((rc.bool() == oc).all()), "max_abs_diff": float((ri[both] - oi[both]).abs().max()),
            "bit_identical": bool(torch.equal(ri[both], oi[both]))}
        gr = torch.randn((n, 3, 3), generator=g).to(dev)
        rb = FastMinv.Fast3x3Minv_backward(gr, ri)
        ob = ops.minv3x3_backward(gr, ri)
        t[f"Fast3x3Minv_backward n={n} (ms/call)"] = {
            "reference": timed(lambda: FastMinv.Fast3x3Minv_backward(gr, ri)),
            "ours": timed(lambda: ops.minv3x3_backward(gr, ri)),
            "max_rel_diff": float(((rb - ob).abs() / rb.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-6)).max())}
        del m, gr, ri, oi, rb, ob
    GS = build_ref.load("GridSamplerMine")
    ws = synth.skinning_voxel((65, 225, 129), seed=7, device=dev)
    ws_cl = ops.voxel_to_channels_last(ws)
    for P in (1 << 20, 1 << 24):
        grid = ((torch.rand((1, 1, 1, P, 3), generator=g) - 0.5) * 2.2).to(dev)
        ro = GS.forward(ws, grid, 0, 1)
        oo = ops.grid_sample3d_forward(ws, grid)
        t[f"GridSamplerMine.forward 24ch 65x225x129, P={P} (ms/call)"] = {
            "reference": timed(lambda: GS.forward(ws, grid, 0, 1), reps=5),
            "ours_same_layout": timed(lambda: ops.grid_sample3d_forward(ws, grid), reps=5),
            "ours_channels_last": timed(lambda: ops.grid_sample3d_forward(
                ws_cl.view(1, *ws_cl.shape), grid, 1), reps=5),
            "bit_identical": bool(torch.equal(ro, oo)), "max_abs_diff": float((ro - oo).abs().max())}
        go = torch.randn_like(ro)
        rgi, rgg = GS.backward(ws, grid, go, 0, 1)
        _, ogg = ops.grid_sample3d_backward(ws, grid, go, need_grad_input=False)
        t[f"GridSamplerMine.backward P={P} (ms/call)"] = {
            "reference": timed(lambda: GS.backward(ws, grid, go, 0, 1), reps=3),
            "ours_frozen_voxel": timed(lambda: ops.grid_sample3d_backward(ws, grid, go, need_grad_input=False), reps=3),
            "grad_grid_max_rel_diff": float((rgg - ogg).abs().max() / rgg.abs().max())}
        del grid, ro, oo, go, rgi, rgg, ogg
    del ws, ws_cl
    I2 = build_ref.load("interp2x_boundary3d")
    x = torch.randn((1, 1, 129, 129, 129), generator=g).to(dev)
    ro, rb = I2.forward(x, 0.0)
    oo, ob = ops.interp2x_boundary3d_forward(x, 0.0, order=0)
    t["interp2x_boundary3d.forward 129^3 -> 257^3 (ms/call)"] = {
        "reference": timed(lambda: I2.forward(x, 0.0)), "ours": timed(lambda: ops.interp2x_boundary3d_forward(x, 0.0, order=0)),
        "values_bit_identical": bool(torch.equal(ro, oo)), "flags_equal": bool(torch.equal(rb.bool(), ob))}
    report["timings_ms"] = t
    json.dump(report, open(os.path.join(OUT, "ref_natives.json"), "w"), indent=1)
    print(json.dumps(report["timings_ms"], indent=1))


if __name__ == "__main__":
    main()
