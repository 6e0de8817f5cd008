"""Where the 0.2 ms of one 257^3 marching-cubes call go: count(+scan+D2H) / allocation / emit, host-timed with a
synchronise after every phase (so launch latencies are included), plus the un-split call for reference."""
import os, sys, time
import torch
from ctypes import byref, c_float, c_int64, c_size_t
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recmv_b200 import _lib, ops, synth
dev = torch.device("cuda", 0)
grid = synth.sphere_sdf_grid(257, num=8, seed=3, device=dev)
lib = _lib.load()
for _ in range(3):
    ops.mc_gpu(grid, 2 / 256, 2 / 256, 2 / 256, -1.0, -1.0, -1.0, 0.0)
nb = c_size_t(0)
lib.recmv_mc_scratch_bytes(257, 257, 257, byref(nb))
scratch = torch.empty((nb.value,), dtype=torch.uint8, device=dev)
V, F = c_int64(0), c_int64(0)
step = (c_float * 3)(2 / 256, 2 / 256, 2 / 256)
org = (c_float * 3)(-1.0, -1.0, -1.0)
t = [0.0, 0.0, 0.0]
reps = 200
torch.cuda.synchronize()
for _ in range(reps):
    a = time.perf_counter()
    lib.recmv_mc_count(grid.data_ptr(), 257, 257, 257, 0.0, scratch.data_ptr(), byref(V), byref(F), None)
    b = time.perf_counter()
    verts = torch.empty((V.value, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((F.value, 3), dtype=torch.int64, device=dev)
    c = time.perf_counter()
    lib.recmv_mc_emit(grid.data_ptr(), 257, 257, 257, 0.0, scratch.data_ptr(), step, org, verts.data_ptr(), faces.data_ptr(), None)
    torch.cuda.synchronize()
    d = time.perf_counter()
    t[0] += b - a; t[1] += c - b; t[2] += d - c
print(f"count+scan+D2H {t[0] / reps * 1e6:.1f} us | alloc {t[1] / reps * 1e6:.1f} us | emit (2 kernels + sync) {t[2] / reps * 1e6:.1f} us | V={V.value} F={F.value}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.mc_gpu(grid, 2 / 256, 2 / 256, 2 / 256, -1.0, -1.0, -1.0, 0.0)
e1.record(); torch.cuda.synchronize()
print(f"whole call: {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
