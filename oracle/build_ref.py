#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- builds the reference's OWN native extensions, unmodified, for sm_100a.

Sources are compiled where they lie under /root/reference (never copied into this repository); the
outputs go to oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the built modules travel to the GPU
box where /root/reference does not exist).  Only tests/, tools/ and bench.py's reference legs import
what is built here; the product (recmv_b200/) never does.

    MCGpu                <- MCGpu/{MCGpu.cpp,CudaKernels.cu}                     (mc_gpu, mc_init)
    FastMinv             <- FastMinv/{M3x3Inv.cpp,Matrix3x3InvKernels.cu}        (Fast3x3Minv[_backward])
    interp2x_boundary3d  <- MCAcc/cuda/interp2x_boundary3d{.cpp,_kernel.cu}
    GridSamplerMine      <- MCAcc/cuda/GridSamplerMine{.cpp,Kernel.cu}; the .cu calls `input.type()` inside
                            AT_DISPATCH_FLOATING_TYPES (lines 931/963/1001), which torch >= 2.1 rejects: the build
                            compiles a sed-patched temporary (`.type()` -> `.scalar_type()`) under /tmp, nothing
                            else changes.

The reference's setup.py files (torch CUDAExtension) are not run: this is the same compile, spelled out
(nvcc -gencode arch=compute_100a,code=sm_100a for the .cu, g++ for the .cpp, one link per module).
"""
import os
import re
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RECMV_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

MODULES = {
    "MCGpu": ("MCGpu", ["MCGpu.cpp", "CudaKernels.cu"], {}),
    "FastMinv": ("FastMinv", ["M3x3Inv.cpp", "Matrix3x3InvKernels.cu"], {}),
    "interp2x_boundary3d": ("MCAcc/cuda", ["interp2x_boundary3d.cpp", "interp2x_boundary3d_kernel.cu"], {}),
    "GridSamplerMine": ("MCAcc/cuda", ["GridSamplerMine.cpp", "GridSamplerMineKernel.cu"],
                        {"GridSamplerMineKernel.cu": (r"\b(input|grad_output)\.type\(\)", r"\1.scalar_type()")}),
}


def build(verbose=False):
    if not os.path.isdir(REF):
        return False
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    incs = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    common = ["-O2", "-std=c++17", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for name, (sub, srcs, patches) in MODULES.items():
        target = os.path.join(OUT, name + ext)
        src_paths = [os.path.join(REF, sub, s) for s in srcs]
        if os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in src_paths) \
                and os.path.getmtime(target) >= os.path.getmtime(__file__):
            continue
        with tempfile.TemporaryDirectory(prefix="recmv_ref_") as tmp:
            objs = []
            for s, sp in zip(srcs, src_paths):
                src = sp
                if s in patches:   # build-time patch of a temporary copy; the include dir stays the reference's
                    pat, rep = patches[s]
                    src = os.path.join(tmp, s)
                    open(src, "w").write(re.sub(pat, rep, open(sp).read()))
                obj = os.path.join(tmp, s + ".o")
                defs = common + [f"-DTORCH_EXTENSION_NAME={name}", f"-I{os.path.join(REF, sub)}"] + incs
                if s.endswith(".cu"):
                    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-w",
                           "--expt-relaxed-constexpr"] + defs + ["-c", src, "-o", obj]
                else:
                    cmd = ["g++", "-fPIC", "-w"] + defs + ["-c", src, "-o", obj]
                if verbose:
                    print(" ".join(cmd))
                subprocess.check_call(cmd)
                objs.append(obj)
            link = ["g++", "-shared", "-o", target] + objs + [f"-L{libdir}", "-L/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu",
                                                               "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
                                                               "-lcudart", f"-Wl,-rpath,{libdir}"]
            subprocess.check_call(link)
    return True


def load(name):
    """Import a built reference module (None if it was never built -- e.g. a fresh clone on the GPU box)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    path = os.path.join(OUT, name + ext)
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", sorted(os.listdir(OUT)) if ok else f"{REF} absent, nothing built")
