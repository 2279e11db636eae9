"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/glic_b200.h declares, sizes behave, and argument validation fails loudly (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "glic_b200.h")).read()
    return sorted(set(re.findall(r"^GLIC_API [^\n(]*?\b(glic_[a-z0-9_]+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from gaussian_lic_b200 import capi
    names = _declared()
    assert len(names) >= 23
    for n in names:
        assert hasattr(capi.lib, n), "libglic_b200.so does not export %s" % n
    assert set(capi.SYMBOLS) == set(names), set(capi.SYMBOLS) ^ set(names)
    assert capi.lib.glic_abi_version() == 1


def test_torch_shim_exports_reference_symbols():
    """The LibTorch shim must define the reference's C++ symbols (rasterize_points.h, ssim.h, spatial.h)."""
    import subprocess
    so = os.path.join(ROOT, "gaussian_lic_b200", "glic_b200_torch.so")
    assert os.path.isfile(so), "build with python gaussian_lic_b200/build.py --torch"
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", so], capture_output=True, text=True).stdout
    for sym in ("RasterizeGaussiansCUDA(", "RasterizeGaussiansBackwardCUDA(", "adamUpdate(", "fusedssim(",
                "fusedssim_backward(", "distCUDA2("):
        assert sym in out, sym


def test_workspace_sizes_and_bucket_bound():
    from gaussian_lic_b200 import capi
    lib = capi.lib
    assert lib.glic_geom_bytes(0) > 0
    assert lib.glic_geom_bytes(1000) >= 1000 * (48 + 4 + 1)
    assert lib.glic_geom_bytes(2000) > lib.glic_geom_bytes(1000)
    assert lib.glic_image_bytes(1920, 1080) >= 1920 * 1080 * 16
    assert lib.glic_binning_bytes(10 ** 6) >= 10 ** 6 * 16          # u32 tile keys + u32 values, ping-pong
    T = 120 * 68
    for R in (0, 1, 31, 32, 33, 5000, 2 * 10 ** 6):
        mb = lib.glic_max_buckets(R, 1920, 1080)
        assert mb == R // 32 + min(T, R)
        assert lib.glic_sample_bytes(R, 1920, 1080) >= mb * (256 * 16 + 4)


def test_bucket_bound_property():
    """B = sum ceil(n_t/32) never exceeds floor(R/32) + min(T, R) (hypothesis)."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    from gaussian_lic_b200 import capi

    @hyp.given(st.lists(st.integers(0, 5000), min_size=1, max_size=12))
    @hyp.settings(max_examples=200, deadline=None)
    def prop(counts):
        W, H = 64, 48  # 4 x 3 = 12 tiles
        counts = (counts + [0] * 12)[:12]
        R = sum(counts)
        B = sum((c + 31) // 32 for c in counts)
        assert B <= capi.lib.glic_max_buckets(R, W, H)

    prop()


def test_argument_validation_without_gpu():
    from gaussian_lic_b200 import capi
    lib = capi.lib
    R = C.c_int64(-5)
    rc = lib.glic_forward_preprocess(10, 3, 15, None, None, 1.0, None, None, None, None, None, 0, None, None, 0, None, 0,
                                     C.byref(R), None)
    assert rc == -1 and b"view" in lib.glic_last_error()
    v = capi.View(1, 1, 1, 1.0, 1.0, -1.0, 1.0, -1.0, 1.0, 64, 48)
    rc = lib.glic_forward_preprocess(10, 7, 15, None, None, 1.0, None, None, None, None, C.byref(v), 0, None, None, 0, None, 0,
                                     C.byref(R), None)
    assert rc == -1 and b"sh_degree" in lib.glic_last_error()
    rc = lib.glic_forward_preprocess(10, 3, 15, None, None, 1.0, None, None, None, None, C.byref(v), 0, None, None, 0, None, 0,
                                     C.byref(R), None)
    assert rc == -2 and b"workspace" in lib.glic_last_error()          # image workspace missing
    rc = lib.glic_forward_preprocess(10, 3, 3, None, None, 1.0, None, None, None, None, C.byref(v), 0, None, None, 0, None, 0,
                                     C.byref(R), None)
    assert rc == -1 and b"coefficients" in lib.glic_last_error()       # degree 3 needs M >= 15
    assert lib.glic_sort_pairs_u64_u32(-1, 45, None, None, None, None, None, 0, None) == -1
    assert lib.glic_sort_pairs_u64_u32(0, 45, None, None, None, None, None, 0, None) == 0
    assert lib.glic_adam_update(None, None, None, None, None, 0.1, 0.9, 0.999, 1e-15, 10, 3, None) == -1
    assert lib.glic_fused_ssim(1, 3, 0, 10, 1e-4, 9e-4, None, None, None, None, None, None, None) == -1
    assert lib.glic_knn_mean_dist2(0, None, None, None, 0, None) == 0
    assert lib.glic_knn_mean_dist2(-1, None, None, None, 0, None) == -1


def test_packed_layout_and_exchange_sizes():
    """Host-only entry points of the packed model step and of the exchange buffer."""
    from gaussian_lic_b200 import capi
    lib = capi.lib
    off = (C.c_size_t * 6)()
    for P, M in ((0, 15), (1, 0), (1000, 15), (500_000, 15)):
        assert lib.glic_packed_offsets(P, M, off) == 0
        k = (4, 3, 3, 1, 3, 3 * M)
        want, acc = [], 0
        for kk in k:
            want.append(acc)
            acc += P * kk
        assert [int(o) for o in off] == want and lib.glic_packed_floats(P, M) == acc
        assert int(off[0]) % 4 == 0                                        # rotations first: float4 stores stay aligned
    assert lib.glic_packed_offsets(10, 15, None) == -1
    # exchange buffer: floats, visibility bytes and the flag block, each padded to 256 B
    n = lib.glic_packed_floats(500_000, 15)
    b = lib.glic_p2p_buffer_bytes(n, 500_000)
    assert b == -(-n * 4 // 256) * 256 + -(-500_000 // 256) * 256 + 256
    assert lib.glic_p2p_buffer_bytes(0, 0) == 256
    # argument validation, nothing launched
    lr = (C.c_float * 6)(*([1e-3] * 6))
    assert lib.glic_adam_update_packed(None, None, None, None, None, lr, 0.9, 0.999, 1e-15, 10, 15, None) == -1
    assert lib.glic_adam_update_packed(None, None, None, None, None, lr, 0.9, 0.999, 1e-15, 0, 15, None) == 0
    assert lib.glic_activations_forward(10, None, None, None, None, None, None, None) == -1
    assert lib.glic_activations_forward(0, None, None, None, None, None, None, None) == 0
    assert lib.glic_activations_backward(-1, None, None, None, None, None, None, None) == -1
    peers = (C.c_void_p * 2)()
    assert lib.glic_p2p_allreduce_mean(2, 2, peers, 10, 10, None) == -1     # rank out of range
    assert lib.glic_p2p_allreduce_mean(0, 9, peers, 10, 10, None) == -1     # more ranks than one NVSwitch domain
    assert lib.glic_p2p_alloc(1024, None, None) == -1


def test_p2p_slices_partition_the_payload():
    """Every 16-byte unit of the gradient block and of the visibility block belongs to exactly one rank's slice."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    from gaussian_lic_b200 import capi
    lib = capi.lib

    @hyp.given(st.integers(0, 3_000_000), st.integers(0, 70_000), st.integers(1, 8))
    @hyp.settings(max_examples=300, deadline=None)
    def prop(n_floats, n_vis, world):
        out = (C.c_size_t * 6)()
        f_end = v_end = 0
        for r in range(world):
            assert lib.glic_p2p_slice(r, world, n_floats, n_vis, out) == 0
            lo, hi, vlo, vhi, f_bytes, flag_off = [int(x) for x in out]
            assert lo == f_end and hi >= lo and vlo == v_end and vhi >= vlo          # contiguous, in rank order
            f_end, v_end = hi, vhi
            assert f_bytes % 256 == 0 and f_bytes >= n_floats * 4 and flag_off % 256 == 0 and flag_off - f_bytes >= n_vis
        assert f_end * 16 == f_bytes and v_end * 16 == flag_off - f_bytes            # ... and complete
        assert flag_off + 256 == lib.glic_p2p_buffer_bytes(n_floats, n_vis)

    prop()
    out = (C.c_size_t * 6)()
    assert lib.glic_p2p_slice(3, 3, 10, 10, out) == -1 and lib.glic_p2p_slice(0, 9, 10, 10, out) == -1


def test_package_fails_loudly_without_library(tmp_path, monkeypatch):
    """No silent CPU/PyTorch fallback: a missing .so is an ImportError."""
    import importlib
    from gaussian_lic_b200 import capi
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        capi._load()
    importlib.reload(capi)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "gaussian_lic_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "glic_oracle" not in src, f


def test_header_is_plain_c_and_links(tmp_path):
    """include/glic_b200.h must be consumable by a C compiler (no C++ or torch types in the signatures) and a C program must
    link against libglic_b200.so and resolve every declared entry point."""
    import subprocess
    from gaussian_lic_b200 import capi
    names = _declared()
    src = tmp_path / "use.c"
    body = "\n".join("    p[%d] = (fn)&%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "glic_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n    fn p[%d];\n%s\n'
                   '    printf("%%d %%d\\n", glic_abi_version(), (int)(sizeof(p) / sizeof(p[0])));\n'
                   '    glic_view v; v.width = 64; v.height = 48; (void)v;\n    return glic_geom_bytes(10) > 0 ? 0 : 1;\n}\n'
                   % (len(names), body))
    exe = tmp_path / "use"
    libdir = os.path.dirname(capi.LIB_PATH)
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:libglic_b200.so", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["1", str(len(names))], (r.stdout, r.stderr)
