"""CPU: FindSurfacePs (fragment decode, pure index logic) against the reference's output (golden) and,
in the container, against the imported reference function."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import refload
from recmv_b200.utils import FindSurfacePs


class Frags:
    def __init__(self, p2f, bary):
        self.pix_to_face, self.bary_coords = p2f, bary


def _inputs():
    g = load_golden("findsurface.npz")
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    return g, t


def test_findsurfaceps_matches_golden():
    g, t = _inputs()
    b, r, c, pts, fi = FindSurfacePs(t["verts"], t["faces"], Frags(t["pix_to_face"], t["bary"]))
    assert np.array_equal(b.numpy(), g["batch"]) and np.array_equal(r.numpy(), g["row"])
    assert np.array_equal(c.numpy(), g["col"]) and np.array_equal(fi.numpy(), g["finds"])  # indices: bit exact
    assert np.allclose(pts.numpy(), g["pts"], atol=1e-6)
    # nothing covered / K = 1
    e = FindSurfacePs(t["verts"], t["faces"], Frags(torch.full((1, 4, 4, 1), -1), torch.rand(1, 4, 4, 1, 3)))
    assert all(x.numel() == 0 for x in e[:3]) and e[3].shape == (0, 3)


@pytest.mark.skipif(not refload.available(), reason="needs /root/reference (container only)")
def test_findsurfaceps_matches_reference_function():
    g, t = _inputs()
    ref = refload.load().FindSurfacePs.FindSurfacePs(t["verts"], t["faces"], Frags(t["pix_to_face"], t["bary"]))
    ours = FindSurfacePs(t["verts"], t["faces"], Frags(t["pix_to_face"], t["bary"]))
    for a, b in zip(ref, ours):
        assert torch.equal(a, b) or torch.allclose(a, b, atol=1e-6)
