#!/usr/bin/env python
"""Link-level drop-in proof: reference host code (unchanged, from /root/reference) + OUR libraries.

Output: oracle/_ref/glic_dropin_ext.so.  Test infrastructure; see dropin_binding.cpp.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GLIC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref", "dropin")
PKG = os.path.join(ROOT, "gaussian_lic_b200")
NAME = "glic_dropin_ext"


def build(verbose=False):
    so = os.path.join(ROOT, "oracle", "_ref", NAME + ".so")
    if not os.path.isdir(os.path.join(REF, "src", "rasterizer")):
        return so if os.path.isfile(so) else None
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ["CXX"] = "/usr/bin/g++"
    from torch.utils import cpp_extension
    forced = ["-include", "cstdint", "-include", "cfloat", "-include", "tuple"]
    cpp_extension.load(
        name=NAME,
        sources=[os.path.join(HERE, "dropin_binding.cpp"), os.path.join(REF, "src", "rasterizer", "rasterizer.cpp")],
        extra_include_paths=[os.path.join(REF, "src")],
        extra_cflags=["-O2", "-std=c++17", "-w"] + forced,
        extra_ldflags=["-L" + PKG, "-l:glic_b200_torch.so", "-lglic_b200", "-Wl,-rpath," + PKG, "-Wl,--no-undefined",
                       "-L" + os.path.join(os.path.dirname(__import__("torch").__file__), "lib"), "-ltorch_python",
                       "-lpython3.12"],
        with_cuda=True, build_directory=OUT, is_python_module=False, verbose=verbose)
    import shutil
    shutil.copy2(os.path.join(OUT, NAME + ".so"), so)
    return so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
