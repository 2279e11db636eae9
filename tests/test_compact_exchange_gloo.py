"""World-size-2 gloo test (CPU) of the identity the mapper's COMPACT exchange relies on.

Two ranks each run the CPU oracle (forward + loss + backward) on their own view of the same Gaussians.  They exchange only
what the mapper exchanges -- an all-gather of the raw dL/dcolour block + the flag byte (radii > 0 | clamp bits) of every view,
and a mean all-reduce of the geometric gradients -- and every rank REBUILDS dL/d(dc, sh-rest) from the gathered colour
gradients with the SH basis of each view's direction.  The result must equal the mean of the full per-view gradients (the
round-1 exchange: all-reduce of all 59 floats), the visibility union must match, and both ranks must hold identical bytes.
The SH basis below restates forward.cu:29-77 on its own (numpy), independent of csrc/sh_math.cuh."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")

C0, C1 = 0.28209479177387814, 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
      -0.5900435899266435]


def sh_rest_basis(d):
    """[P,3] unit directions -> [P,15] basis values of the degree 1..3 real spherical harmonics (forward.cu:29-77)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    return np.stack([-C1 * y, C1 * z, -C1 * x,
                     C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy),
                     C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy), C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
                     C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy)], 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    from gaussian_lic_b200 import synthetic as syn
    from helpers import small_scene
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, W, H, deg = 1500, 160, 112, 3
    g, cam = small_scene(P, W, H, 31, deg, view=(0, 3)[rank])              # same Gaussians, this rank's view
    o = Oracle(np.float64)
    f = o.forward(g, cam)
    _, dl = o.loss(f["color"], syn.make_gt_image(W, H, seed=5 + rank))
    b = o.backward(f, dl)
    clamped = o.state(f)["clamped"]
    o.free(f)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float64))
    # ---- round-1 exchange: mean of the full gradients (the reference semantics of the batch) ----
    full = {k: t(b[k]).clone() for k in ("dL_ddc", "dL_dsh", "dL_dmeans3D", "dL_dscales", "dL_drots", "dL_dopacity")}
    for v in full.values():
        dist.all_reduce(v)
        v /= world
    # ---- compact exchange: geometric mean-reduce + all-gather of (dL/dcolour, flags, camera centre) ----
    geo = {k: t(b[k]).clone() for k in ("dL_dmeans3D", "dL_dscales", "dL_drots", "dL_dopacity")}
    for v in geo.values():
        dist.all_reduce(v)
        v /= world
    flags = ((f["radii"] > 0).astype(np.uint8) << 7) | (clamped[:, 0] | (clamped[:, 1] << 1) | (clamped[:, 2] << 2)).astype(np.uint8)
    mine = (t(b["dL_dcolors"]), torch.as_tensor(flags), t(cam["campos"]))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    g_dc = np.zeros((P, 3)); g_sh = np.zeros((P, 15, 3)); vis = np.zeros(P, bool)
    for col, fl, campos in gathered:                                       # slot order: identical on every rank
        col, fl = col.numpy(), fl.numpy()
        seen = (fl & 0x80) != 0
        mask = np.stack([(fl >> c) & 1 for c in range(3)], 1) == 0         # clamped channels carry no gradient
        gm = np.where(seen[:, None] & mask, col, 0.0)
        d = g["means"].astype(np.float64) - campos.numpy()
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        g_dc += C0 * gm
        g_sh += sh_rest_basis(d)[:, :, None] * gm[:, None, :]
        vis |= seen
    g_dc /= world; g_sh /= world
    torch.save(dict(full={k: v.numpy() for k, v in full.items()}, geo={k: v.numpy() for k, v in geo.items()}, g_dc=g_dc, g_sh=g_sh,
                    vis=vis, radii=f["radii"].copy()), os.path.join(out_dir, "c%d.pt" % rank))
    dist.destroy_process_group()


def test_compact_exchange_equals_full_gradient_mean(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(tmp_path, "c%d.pt" % i), weights_only=False) for i in range(2)]
    for i in range(2):
        scale = np.abs(r[i]["full"]["dL_dsh"]).max()
        assert scale > 0
        np.testing.assert_allclose(r[i]["g_sh"], r[i]["full"]["dL_dsh"], rtol=0, atol=1e-12 * max(scale, 1.0))
        np.testing.assert_allclose(r[i]["g_dc"], r[i]["full"]["dL_ddc"].reshape(-1, 3), rtol=0, atol=1e-12)
        for k in r[i]["geo"]:
            np.testing.assert_array_equal(r[i]["geo"][k], r[i]["full"][k])
    union = (r[0]["radii"] > 0) | (r[1]["radii"] > 0)
    assert np.array_equal(r[0]["vis"], union) and np.array_equal(r[1]["vis"], union) and union.sum() > 100
    for k in ("g_dc", "g_sh"):                                              # replicas rebuild the same bytes
        assert np.array_equal(r[0][k], r[1][k])
