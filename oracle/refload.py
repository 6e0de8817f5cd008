"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (recmv_b200/).

Stub-import loader for the *unmodified* reference Python modules under
/root/reference (model/network.py, model/Deformer.py, model/RenderNet.py,
model/Embedder.py, utils/*.py, MCAcc/seg3d_lossless.py).

The reference's packages import third-party modules that are absent here
(pytorch3d, torch_scatter, openmesh, trimesh, smpl_pytorch, ...) and its three
CUDA extensions (FastMinv, MCGpu, GridSamplerMine).  This loader
  * appends a meta-path finder that satisfies any import that *originates from
    a file under /root/reference* and cannot be resolved with an inert stub
    module, and
  * injects three semantic shims under the reference's own names:
      GridSamplerMine.{forward,backward,dbackward}  -> oracle_torch.grid_sample3d_*
          (reference asserts forward == F.grid_sample(bilinear,border,
           align_corners=False): MCAcc/check_grid_sampler_mine.py:8-9)
      FastMinv.Fast3x3Minv[_backward]              -> oracle_torch.minv3x3_*
          (restating FastMinv/Matrix3x3InvKernels.cu:21-104)
      smpl_pytorch.util.batch_rodrigues            -> oracle_torch.batch_rodrigues
          (un-vendored dependency; parity UNPINNED -- bone matrices are inputs
           to every kernel so no kernel's parity depends on it)

It only works in the build container (where /root/reference exists); the GPU box
uses the committed fixtures under tests/golden/ instead.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("RECMV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


class _Stub(types.ModuleType):
    """Inert module: any attribute is another stub / a callable returning None."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _StubAttr(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child


class _StubAttr:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _StubAttr(f"{self._name}.{name}")

    def __mro_entries__(self, bases):  # allows `class X(stub.Base)`
        return (object,)

    def __iter__(self):
        return iter(())


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []  # behave like a package so submodule imports resolve
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    """Last-resort finder: only answers for imports issued from reference files."""

    def find_spec(self, fullname, path, target=None):
        f = sys._getframe(1)
        from_ref = False
        depth = 0
        while f is not None and depth < 40:
            fn = f.f_code.co_filename
            if fn.startswith(REF_ROOT):
                from_ref = True
                break
            f = f.f_back
            depth += 1
        root = fullname.split(".")[0]
        if not from_ref and root not in _STUBBED_ROOTS:
            return None
        _STUBBED_ROOTS.add(root)
        return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)


_STUBBED_ROOTS = set()
_installed = False


def install():
    """Make `import model.network`, `import utils`, `import MCAcc...` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    from . import oracle_torch as ot

    # --- semantic shims under the reference's names -------------------------------------
    gsm = types.ModuleType("GridSamplerMine")
    gsm.forward = lambda inp, grid, interp, pad: ot.grid_sample3d_fwd(inp, grid)
    gsm.backward = lambda inp, grid, gout, interp, pad: ot.grid_sample3d_bwd(inp, grid, gout)
    gsm.dbackward = lambda ggi, ggg, inp, grid, gout, interp, pad: ot.grid_sample3d_bwd2(
        ggi, ggg, inp, grid, gout)
    sys.modules["GridSamplerMine"] = gsm

    fm = types.ModuleType("FastMinv")
    fm.Fast3x3Minv = lambda ms: list(ot.minv3x3_fwd(ms))
    fm.Fast3x3Minv_backward = lambda g, inv: ot.minv3x3_bwd(g, inv)
    sys.modules["FastMinv"] = fm

    smpl = _Stub("smpl_pytorch")
    smpl.__path__ = []
    smpl_util = _Stub("smpl_pytorch.util")
    smpl_util.batch_rodrigues = ot.batch_rodrigues
    smpl.util = smpl_util
    sys.modules["smpl_pytorch"] = smpl
    sys.modules["smpl_pytorch.util"] = smpl_util
    _STUBBED_ROOTS.add("smpl_pytorch")

    ts = types.ModuleType("torch_scatter")
    ts.scatter = ot.scatter
    sys.modules["torch_scatter"] = ts

    sys.meta_path.append(_StubFinder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def load():
    """Returns a namespace with the reference classes used by the golden generator."""
    install()
    ns = types.SimpleNamespace()
    ns.network = importlib.import_module("model.network")
    ns.Deformer = importlib.import_module("model.Deformer")
    ns.RenderNet = importlib.import_module("model.RenderNet")
    ns.Embedder = importlib.import_module("model.Embedder")
    ns.utils = importlib.import_module("utils")
    ns.utils_utils = importlib.import_module("utils.utils")
    ns.FindSurfacePs = importlib.import_module("utils.FindSurfacePs")
    ns.MCAcc = importlib.import_module("MCAcc")
    return ns
