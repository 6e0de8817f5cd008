"""Marching cubes pinned to the REFERENCE'S OWN EXECUTION: tests/golden/mc_ref.npz holds the output of the
reference's unmodified MCGpu extension (MCGpu/{MCGpu.cpp,CudaKernels.cu} built for sm_100a by oracle/build_ref.py,
run on a B200 by tools/dump_ref_natives.py) in canonical order (the reference emits through atomics), for four
small grids in full and -- counts, SHA-256 of the canonical face array, float64 vertex sums and a 4096-row sample --
for the three anisotropic production pyramids of train.py:47-71 and the 257^3 benchmark grid.

CPU: the C restatement oracle/mc_oracle.c (which the parity tests at other sizes use) reproduces the faces bit for bit
and the vertices to 1e-6 (it is compiled without FMA contraction).
GPU: ops.mc_gpu (through the C ABI) reproduces them: faces bit exact, vertices bit identical."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mc_oracle
from recmv_b200 import synth

SMALL = ["s41", "a21x37x13", "a33x17x50", "s41_iso0.1"]
LARGE = ["coarse225x321x129", "medium289x385x193", "fine321x417x225", "s257"]


def _grid(g, name, device="cpu"):
    shape = tuple(int(v) for v in g[name + "_shape"])
    num, seed = (int(v) for v in g[name + "_meta"])
    sdf = synth.sphere_sdf_grid(shape, num=num, seed=seed, device=device)
    step = tuple(2.0 / (n - 1) for n in shape)
    return sdf, step, float(g[name + "_iso"][0])


def _check(g, name, verts, faces, exact_verts):
    cv, cf = mc_oracle.canonical(verts, faces)
    V, F = (int(v) for v in g[name + "_counts"])
    assert cv.shape[0] == V and cf.shape[0] == F
    sha = hashlib.sha256(np.ascontiguousarray(cf.astype(np.int64)).tobytes()).digest()
    assert sha == bytes(g[name + "_faces_sha256"]), f"{name}: canonical faces differ from the reference kernel's"
    assert np.allclose(cv.astype(np.float64).sum(0), g[name + "_verts_sum"], rtol=0, atol=1e-6 * max(V, 1))
    if name + "_verts" in g:
        assert np.array_equal(cf, g[name + "_faces"])
        ref = g[name + "_verts"]
    else:
        assert np.array_equal(cf[g[name + "_faces_idx"]], g[name + "_faces_sample"])
        cv, ref = cv[g[name + "_verts_idx"]], g[name + "_verts_sample"]
    if exact_verts:
        assert np.array_equal(cv, ref)                         # bit identical vertex coordinates
    else:
        assert np.abs(cv - ref).max() < 1e-6                   # the C restatement is built without FMA contraction


@pytest.mark.parametrize("name", SMALL + ["coarse225x321x129"])
def test_c_restatement_reproduces_the_reference_kernel(name):
    """Pins oracle/mc_oracle.c on the reference's own output (the grid is rebuilt on the CPU: torch's CPU and CUDA
    elementwise sqrt/sub/min are correctly rounded, so it is the same grid the reference kernel saw)."""
    g = load_golden("mc_ref.npz")
    sdf, step, iso = _grid(g, name)
    v, f = mc_oracle.marching_cubes(sdf.numpy(), step, (-1.0, -1.0, -1.0), iso)
    _check(g, name, v, f, exact_verts=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL + LARGE)
def test_mc_gpu_reproduces_the_reference_kernel(name):
    from recmv_b200 import ops
    g = load_golden("mc_ref.npz")
    sdf, step, iso = _grid(g, name, "cuda:0")
    v, f = ops.mc_gpu(sdf, *step, -1.0, -1.0, -1.0, iso)
    assert v.dtype == torch.float32 and f.dtype == torch.int64
    _check(g, name, v.cpu().numpy(), f.cpu().numpy(), exact_verts=True)
