"""Drop-in for the reference's `MCGpu` extension module (MCGpu/MCGpu.cpp:58-61)."""
from recmv_b200.ops import mc_gpu  # noqa: F401


def mc_init(device_id):
    """MCGpu.mc_init: the reference pre-creates its per-device singleton; nothing to do here
    (scratch is grow-only and allocated on first use)."""
    return None
