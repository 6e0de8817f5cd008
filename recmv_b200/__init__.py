"""recmv_b200 -- B200-native (sm_100a) implementation of REC-MV's implicit-surface hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/recmv_b200.h) -> librecmv_b200.so
  _lib.py, ops.py  ctypes binding and torch-facing wrappers (allocation, streams, autograd wiring)
  model/           mirror of the reference's model.{network,Embedder,Deformer,RenderNet} API
  compat/          drop-in modules named FastMinv / MCGpu / GridSamplerMine (put on sys.path)
  render.py        the fused ray -> sdf render path and its multi-GPU sharding
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
