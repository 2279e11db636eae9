"""gaussian_lic_b200 -- B200-native (sm_100a) differentiable Gaussian rasterizer hot path behind
Gaussian-LIC's operator surface.

Layout: csrc/ (hand-written CUDA + the C ABI + the LibTorch shim), capi.py (ctypes plumbing),
ops.py (host-side mirror of the reference operator interface), synthetic.py (deterministic scenes),
dist.py (view-sharded data parallelism).  The compute path is libglic_b200.so only; importing this
package fails loudly when that library has not been built.
"""
from . import capi  # noqa: F401  (raises ImportError if libglic_b200.so is missing)
from . import synthetic  # noqa: F401

__all__ = ["capi", "synthetic"]
