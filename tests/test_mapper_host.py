"""Host-side pieces of the native mapper that need no GPU: the camera block (camera.h:38-110 in closed form) against the
numpy restatement used by every parity test, and the loud failure of the mapper without a CUDA device (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("view", [0, 1, 3, 5, 7])
def test_camera_block_matches_numpy_restatement(view):
    from gaussian_lic_b200 import mapper, synthetic as syn
    W, H, fx, fy, cx, cy = 640, 480, 400.0, 410.0, 330.5, 235.25
    R, t = syn.orbit_pose(view, radius=2.0)
    a = mapper.camera_block(W, H, fx, fy, cx, cy, R, t)
    b = syn.make_camera(W, H, fx, fy, cx, cy, R, t)
    np.testing.assert_array_equal(a["view"], b["view"])                       # Rt, float-rounded from the same doubles
    np.testing.assert_allclose(a["proj"], b["proj"], rtol=0, atol=1e-6)       # 4x4 float product, accumulation order may differ
    np.testing.assert_allclose(a["campos"], b["campos"], rtol=0, atol=5e-7)
    np.testing.assert_array_equal(a["lims"], b["lims"])                       # four DISTINCT clamp limits (cx != W/2, cy != H/2)
    assert len(set(np.round(a["lims"], 6))) == 4
    assert abs(a["tanfovx"] - b["tanfovx"]) <= 1.2e-7 and abs(a["tanfovy"] - b["tanfovy"]) <= 1.2e-7


def test_mapper_refuses_to_run_without_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gaussian_lic_b200 import capi, mapper
    with pytest.raises(capi.GlicError):
        mapper.Mapper(64, 48, 50.0, 50.0, 32.0, 24.0, sh_degree=0, capacity=256)
    bad = capi.MapperConfig()                                                # a zeroed configuration is rejected before any CUDA call
    h = C.c_void_p()
    assert capi.lib.glic_mapper_create(C.byref(bad), C.byref(h)) == -1 and not h.value


def _build_c_host(tmp_path):
    import os
    import subprocess
    from gaussian_lic_b200 import capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "mapper_loop"
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["/usr/bin/gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
           os.path.join(root, "examples", "mapper_loop.c"), "-o", str(exe), "-L", libdir, "-l:libglic_b200.so", "-Wl,-rpath," + libdir, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_host_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/mapper_loop.c: the whole mapping loop from C99 against include/glic_b200.h -- no torch, no Python."""
    import subprocess
    torch = pytest.importorskip("torch")
    exe = _build_c_host(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the run itself is tests/test_gpu_mapper.py::test_plain_c_host_runs_the_loop")
    r = subprocess.run([str(exe), "2", "2000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)
