"""Drop-in for the reference's `FastMinv` extension module (FastMinv/M3x3Inv.cpp:61-64).
Put recmv_b200/compat on sys.path ahead of the reference's build to use it unchanged:
`from FastMinv import Fast3x3Minv, Fast3x3Minv_backward` (utils/utils.py:4)."""
from recmv_b200.ops import minv3x3 as Fast3x3Minv  # noqa: F401
from recmv_b200.ops import minv3x3_backward as Fast3x3Minv_backward  # noqa: F401
