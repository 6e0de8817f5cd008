"""Mirror of the reference's `MCAcc` package surface (MCAcc/__init__.py:1-3)."""
from .seg3d_lossless import Seg3dLossless, create_grid3D  # noqa: F401
from ..ops import GridSamplerMine3dFunction  # noqa: F401
