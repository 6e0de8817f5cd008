"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/mc_oracle.c (built by oracle/Makefile)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmc_oracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def marching_cubes(sdf, step=(1., 1., 1.), origin=(0., 0., 0.), iso=0.0):
    """sdf [NX,NY,NZ] float32 (z fastest) -> (verts [V,3] f32, faces [F,3] i64), sequential order."""
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    sdf = np.ascontiguousarray(sdf, dtype=np.float32)
    NX, NY, NZ = sdf.shape
    step = np.asarray(step, dtype=np.float32)
    origin = np.asarray(origin, dtype=np.float32)
    nv, nf = ctypes.c_longlong(0), ctypes.c_longlong(0)
    cap_v, cap_f = 1 << 16, 1 << 17
    while True:
        verts = np.empty((cap_v, 3), np.float32)
        faces = np.empty((cap_f, 3), np.int64)
        rc = lib.mc_oracle(sdf.ctypes.data_as(ctypes.c_void_p), NX, NY, NZ, ctypes.c_float(iso),
                           step.ctypes.data_as(ctypes.c_void_p), origin.ctypes.data_as(ctypes.c_void_p),
                           verts.ctypes.data_as(ctypes.c_void_p), faces.ctypes.data_as(ctypes.c_void_p),
                           ctypes.c_longlong(cap_v), ctypes.c_longlong(cap_f), ctypes.byref(nv),
                           ctypes.byref(nf))
        if rc == 0:
            return verts[:nv.value].copy(), faces[:nf.value].copy()
        if rc != 1:
            raise MemoryError("mc_oracle")
        cap_v, cap_f = max(cap_v, nv.value), max(cap_f, nf.value)


def canonical(verts, faces):
    """Order-independent form: vertices sorted lexicographically, faces remapped, each face rotated so
    its smallest index leads (winding preserved), faces sorted."""
    order = np.lexsort((verts[:, 2], verts[:, 1], verts[:, 0]))
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order))
    v = verts[order]
    f = np.where(faces >= 0, inv[np.clip(faces, 0, None)], -1)
    k = np.argmin(f, axis=1)
    f = np.stack([np.roll(row, -s) for row, s in zip(f, k)]) if len(f) else f
    f = f[np.lexsort((f[:, 2], f[:, 1], f[:, 0]))] if len(f) else f
    return v, f
