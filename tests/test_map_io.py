"""On-disk map format, CPU only.  glic_ply_write must produce byte-for-byte the file the REFERENCE's own
PLY writer produces for the same tensors: tests/golden/map_*.ply were written by /root/reference/src/tinyply.h driven like
GaussianModel::saveMap (tests/golden/make_golden_ply.py, oracle/ref_build/ply_ref.cpp).  Plus a numpy restatement of the
layout, round trips, and loud failures on foreign files."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _write(lib, path, m, P, M):
    from gaussian_lic_b200 import capi
    a = {k: np.ascontiguousarray(m[k], np.float32) for k in ("xyz", "f_dc", "f_rest", "opacity", "scale", "rotation")}
    capi.check(lib.glic_ply_write(path.encode(), P, M, _ptr(a["xyz"]), _ptr(a["f_dc"]), _ptr(a["f_rest"]), _ptr(a["opacity"]),
                                  _ptr(a["scale"]), _ptr(a["rotation"])), "ply_write")


def _numpy_ply(m, P, M):
    """Independent restatement: header text + row-interleaved float32 body, f_rest channel-major."""
    names = ["x", "y", "z"] + ["f_dc_%d" % i for i in range(3)] + ["f_rest_%d" % i for i in range(3 * M)] + ["opacity"] + \
            ["scale_%d" % i for i in range(3)] + ["rot_%d" % i for i in range(4)]
    head = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P + "".join("property float %s\n" % n for n in names) + "end_header\n"
    rows = np.concatenate([m["xyz"].reshape(P, 3), m["f_dc"].reshape(P, 1, 3).transpose(0, 2, 1).reshape(P, 3),
                           m["f_rest"].reshape(P, M, 3).transpose(0, 2, 1).reshape(P, 3 * M), m["opacity"].reshape(P, 1),
                           m["scale"].reshape(P, 3), m["rotation"].reshape(P, 4)], axis=1).astype("<f4")
    return head.encode() + rows.tobytes()


@pytest.mark.parametrize("name", ["map_deg3", "map_deg0", "map_empty"])
def test_writer_is_byte_identical_to_the_reference_writer(name, tmp_path):
    from gaussian_lic_b200 import capi
    lib = capi.lib
    z = np.load(os.path.join(GOLD, name + "_inputs.npz"))
    P, M = int(z["P"]), int(z["M"])
    m = {k: z[k] for k in ("xyz", "f_dc", "f_rest", "opacity", "scale", "rotation")}
    out = str(tmp_path / "mine.ply")
    _write(lib, out, m, P, M)
    mine, gold = open(out, "rb").read(), open(os.path.join(GOLD, name + ".ply"), "rb").read()
    assert mine == gold, "glic_ply_write differs from the reference's tinyply output"
    assert mine == _numpy_ply(m, P, M)
    assert lib.glic_ply_bytes(P, M) == len(gold)


def test_round_trip_and_packed_writer(tmp_path):
    from gaussian_lic_b200 import capi
    lib = capi.lib
    rng = np.random.default_rng(0)
    P, M = 9001, 15                                           # more than one 4096-row chunk
    m = dict(xyz=rng.normal(size=(P, 3)), f_dc=rng.normal(size=(P, 1, 3)), f_rest=rng.normal(size=(P, M, 3)),
             opacity=rng.normal(size=P), scale=rng.normal(size=(P, 3)), rotation=rng.normal(size=(P, 4)))
    m = {k: v.astype(np.float32) for k, v in m.items()}
    path = str(tmp_path / "map.ply")
    _write(lib, path, m, P, M)
    assert open(path, "rb").read() == _numpy_ply(m, P, M)
    p, mm, off = C.c_uint32(), C.c_uint32(), C.c_size_t()
    capi.check(lib.glic_ply_read_header(path.encode(), C.byref(p), C.byref(mm), C.byref(off)), "hdr")
    assert (p.value, mm.value) == (P, M) and off.value == os.path.getsize(path) - P * 59 * 4
    back = {k: np.zeros_like(v) for k, v in m.items()}
    capi.check(lib.glic_ply_read(path.encode(), P, M, _ptr(back["xyz"]), _ptr(back["f_dc"]), _ptr(back["f_rest"]),
                                 _ptr(back["opacity"]), _ptr(back["scale"]), _ptr(back["rotation"])), "read")
    for k in m:
        assert np.array_equal(back[k], m[k]), k
    # the packed model buffer (rotation | xyz | scale | opacity | dc | rest) writes the same file
    packed = np.concatenate([m["rotation"].ravel(), m["xyz"].ravel(), m["scale"].ravel(), m["opacity"].ravel(),
                             m["f_dc"].ravel(), m["f_rest"].ravel()]).astype(np.float32)
    p2 = str(tmp_path / "packed.ply")
    capi.check(lib.glic_ply_write_packed(p2.encode(), P, M, _ptr(packed)), "packed")
    assert open(p2, "rb").read() == open(path, "rb").read()


def test_reader_fails_loudly_on_foreign_files(tmp_path):
    from gaussian_lic_b200 import capi
    lib = capi.lib
    p, mm = C.c_uint32(), C.c_uint32()
    assert lib.glic_ply_read_header(str(tmp_path / "missing.ply").encode(), C.byref(p), C.byref(mm), None) == -1
    bad = tmp_path / "ascii.ply"
    bad.write_text("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0\n")
    assert lib.glic_ply_read_header(str(bad).encode(), C.byref(p), C.byref(mm), None) == -1
    assert b"unsupported" in lib.glic_last_error() or b"not a binary" in lib.glic_last_error()
    # right format, wrong property order (normals inserted like stock 3DGS): refused, not mis-parsed
    z = np.load(os.path.join(GOLD, "map_deg0_inputs.npz"))
    wrong = tmp_path / "normals.ply"
    head = open(os.path.join(GOLD, "map_deg0.ply"), "rb").read().split(b"end_header\n")[0].decode()
    head = head.replace("property float z\n", "property float z\nproperty float nx\nproperty float ny\nproperty float nz\n")
    wrong.write_bytes(head.encode() + b"end_header\n" + b"\0" * (int(z["P"]) * 20 * 4))
    assert lib.glic_ply_read_header(str(wrong).encode(), C.byref(p), C.byref(mm), None) == -1
    assert b"differ" in lib.glic_last_error()
    assert lib.glic_ply_write(None, 0, 0, None, None, None, None, None, None) == -1
    # truncated body
    trunc = tmp_path / "trunc.ply"
    trunc.write_bytes(open(os.path.join(GOLD, "map_deg3.ply"), "rb").read()[:-8])
    z3 = np.load(os.path.join(GOLD, "map_deg3_inputs.npz"))
    bufs = [np.zeros_like(z3[k], dtype=np.float32) for k in ("xyz", "f_dc", "f_rest", "opacity", "scale", "rotation")]
    assert lib.glic_ply_read(str(trunc).encode(), int(z3["P"]), int(z3["M"]), *[_ptr(b) for b in bufs]) == -1
    assert b"truncated" in lib.glic_last_error()


def test_packed_model_save_and_load_map(tmp_path):
    """PackedModel.save_map / load_map on a CPU-resident model: same bytes as the six-array writer, lossless reload."""
    torch = pytest.importorskip("torch")
    from gaussian_lic_b200 import capi, model
    rng = np.random.default_rng(5)
    P, M = 321, 15
    g = dict(means=rng.normal(size=(P, 3)), log_scales=rng.normal(size=(P, 3)), rots=rng.normal(size=(P, 4)),
             opacity_logits=rng.normal(size=P), dc=rng.normal(size=(P, 3)), sh=rng.normal(size=(P, M, 3)), degree=3)
    g = {k: (v.astype(np.float32) if hasattr(v, "astype") else v) for k, v in g.items()}
    mdl = model.PackedModel(g, torch.device("cpu"))
    path = str(tmp_path / "point_cloud.ply")
    mdl.save_map(path)
    six = str(tmp_path / "six.ply")
    _write(capi.lib, six, dict(xyz=g["means"], f_dc=g["dc"], f_rest=g["sh"], opacity=g["opacity_logits"], scale=g["log_scales"],
                               rotation=g["rots"]), P, M)
    assert open(path, "rb").read() == open(six, "rb").read()
    back = model.PackedModel.load_map(path, torch.device("cpu"))
    assert back.P == P and back.M == M and back.degree == 3
    assert torch.equal(back.params, mdl.params)
    assert float(back.exp_avg.abs().max()) == 0.0
