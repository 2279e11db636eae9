#!/usr/bin/env python
"""Build "Oracle A": the reference's OWN hot-path sources, compiled unmodified for sm_100a.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/README.md).  Sources are compiled from where
they lie under /root/reference (never copied into this repo); the only output is
oracle/_ref/glic_ref_ext.so (git-ignored, NOT gpurun-ignored, so it travels to the GPU box).

Recipe:
  * nvcc 12.9, -std=c++17 -O3, -gencode arch=compute_100a,code=sm_100a (default -fmad=true,
    no fast-math: identical to what the reference's CMakeLists.txt asks for, minus the arch);
  * forced includes <cstdint> <cfloat> <tuple> (missing in the reference under GCC 13);
  * oracle/ref_build/glm/glm.hpp stands in for the un-vendored GLM dependency.

The reference's own build system (catkin/CMake + ROS/OpenCV/Eigen/PCL) is NOT run and could
not be: none of those libraries exist in this image.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GLIC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
NAME = "glic_ref_ext"


def sources():
    s = os.path.join(REF, "src")
    return [
        os.path.join(HERE, "ref_binding.cpp"),
        os.path.join(HERE, "cub_sort.cu"),
        os.path.join(s, "rasterizer", "rasterizer.cpp"),
        os.path.join(s, "rasterizer", "rasterize_points.cu"),
        os.path.join(s, "rasterizer", "cuda_rasterizer", "forward.cu"),
        os.path.join(s, "rasterizer", "cuda_rasterizer", "backward.cu"),
        os.path.join(s, "rasterizer", "cuda_rasterizer", "rasterizer_impl.cu"),
        os.path.join(s, "rasterizer", "cuda_rasterizer", "adam.cu"),
        os.path.join(s, "fused-ssim", "ssim.cu"),
        os.path.join(s, "simple-knn", "simple_knn.cu"),
        os.path.join(s, "simple-knn", "spatial.cu"),
    ]


def is_built():
    return os.path.isfile(os.path.join(OUT, NAME + ".so"))


def build(verbose=False):
    """Returns the path of the built extension, or None when /root/reference is absent."""
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.isdir(os.path.join(REF, "src", "rasterizer")):
        return so if os.path.isfile(so) else None
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils import cpp_extension

    forced = ["-include", "cstdint", "-include", "cfloat", "-include", "tuple"]
    cpp_extension.load(
        name=NAME,
        sources=sources(),
        extra_include_paths=[HERE, os.path.join(REF, "src"), os.path.join(REF, "src", "simple-knn")],
        extra_cflags=["-O3", "-std=c++17", "-w"] + forced,
        extra_cuda_cflags=["-O3", "-std=c++17", "-w", "-lineinfo"] + forced,
        build_directory=OUT,
        with_cuda=True,
        is_python_module=False,
        verbose=verbose,
    )
    return so


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", p)
