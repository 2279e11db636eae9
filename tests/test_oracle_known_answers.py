"""Known-answer and property tests of the CPU oracle -- no GPU needed."""

import numpy as np
import pytest

from gaussian_lic_b200 import synthetic as syn


def _one_gaussian(opacity, xyz=(0.0, 0.0, 5.0), scale=0.05, W=64, H=48, f=100.0):
    g = dict(means=np.array([xyz], np.float32), scales=np.full((1, 3), scale, np.float32),
             rots=np.array([[1, 0, 0, 0]], np.float32), opacity=np.array([opacity], np.float32),
             dc=np.array([[1.0, 0.5, -0.2]], np.float32), sh=np.zeros((1, 15, 3), np.float32), degree=0)
    cam = syn.make_camera(W, H, f, f, W / 2.0, H / 2.0)
    return g, cam


def test_single_isotropic_gaussian_centre_alpha(oracle32):
    """A Gaussian centred exactly on a pixel: alpha = min(0.99, o), C = alpha * max(0, SH_C0*dc + 0.5)."""
    # centre projects to ((0+1)*64-1)/2 = 31.5 -> shift by half a pixel so it lands on pixel 32
    g, cam = _one_gaussian(0.6, xyz=(0.5 * 5.0 / 100.0, 0.5 * 5.0 / 100.0, 5.0))
    f = oracle32.forward(g, cam)
    st = oracle32.state(f)
    np.testing.assert_allclose(st["xy"][0], [32.0, 24.0], atol=1e-4)
    col = 0.28209479177387814 * g["dc"][0] + 0.5
    np.testing.assert_allclose(f["color"][:, 24, 32], 0.6 * np.maximum(col, 0), rtol=1e-5)
    np.testing.assert_allclose(f["final_T"][24, 32], 0.4, rtol=1e-6)
    assert st["clamped"][0].tolist() == [0, 0, 0]
    g2, _ = _one_gaussian(1.0, xyz=g["means"][0])
    f2 = oracle32.forward(g2, cam)
    np.testing.assert_allclose(f2["final_T"][24, 32], 0.01, rtol=1e-4)      # alpha clamped to 0.99
    oracle32.free(f); oracle32.free(f2)


def test_culls(oracle32):
    g, cam = _one_gaussian(0.5, xyz=(0, 0, 0.19))
    f = oracle32.forward(g, cam); assert f["R"] == 0 and f["radii"][0] == 0; oracle32.free(f)      # z <= 0.2
    g, cam = _one_gaussian(1.0 / 256.0)
    f = oracle32.forward(g, cam); assert f["R"] == 0 and f["radii"][0] == 0; oracle32.free(f)      # opacity < 1/255
    g, cam = _one_gaussian(0.5, xyz=(50.0, 0, 5.0))
    f = oracle32.forward(g, cam); assert f["R"] == 0; oracle32.free(f)                              # off screen
    assert not f["color"].any() and np.all(f["final_T"] == 1.0)


def test_two_gaussians_depth_order(oracle32):
    g, cam = _one_gaussian(0.5)
    g = {k: (np.concatenate([v, v]) if isinstance(v, np.ndarray) else v) for k, v in g.items()}
    g["means"][0] = (0.025, 0.025, 6.0); g["means"][1] = (0.02, 0.02, 4.0)     # index 1 is nearer
    g["dc"][0] = (2.0, 2.0, 2.0); g["dc"][1] = (-1.0, -1.0, -1.0)
    f = oracle32.forward(g, cam)
    st = oracle32.state(f)
    r0, r1 = st["ranges"][24 // 16 * 4 + 32 // 16]
    assert st["point_list"][r0:r1].tolist() == [1, 0]                          # sorted by depth, not by index
    oracle32.free(f)


def test_contract_properties(oracle32):
    g, cam = syn.make_scene("cfg1", P=4000)
    f = oracle32.forward(g, cam)
    st = oracle32.state(f)
    assert st["tiles_touched"].sum() == f["R"]
    k = st["keys_sorted"]
    assert np.all(k[1:] >= k[:-1])
    n = (st["ranges"][:, 1] - st["ranges"][:, 0]).astype(np.int64)
    assert n.sum() == f["R"]
    assert ((n + 31) // 32).sum() == f["B"] == st["bucket_offsets"][-1]
    assert f["final_T"].min() >= 1e-4 and f["final_T"].max() <= 1.0
    assert ((f["radii"] > 0) == (st["tiles_touched"] > 0)).all()
    # ties inside a tile keep ascending Gaussian index (stable sort)
    same = k[1:] == k[:-1]
    assert np.all(st["point_list"][1:][same] > st["point_list"][:-1][same])
    oracle32.free(f)


def test_higher_msb_and_sort(oracle32):
    assert oracle32.higher_msb(8160) == 13 and oracle32.higher_msb(1200) == 11 and oracle32.higher_msb(1) == 1
    keys, vals = syn.make_sort_pairs(50_000)
    ko, vo = oracle32.sort_pairs(keys, vals, 45)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])


def test_ssim_identity_and_padding(oracle32):
    rng = np.random.default_rng(0)
    x = rng.random((3, 40, 50), dtype=np.float32)
    m, *_ = oracle32.ssim(x, x)
    np.testing.assert_allclose(m, 1.0, atol=1e-5)                             # SSIM(x, x) = 1 everywhere (zero padding included)
    ones = np.ones((1, 30, 30), np.float32)
    m, *_ = oracle32.ssim(ones, 0.5 * ones)
    mu1 = 1.0; mu2 = 0.5                                                       # interior: sigma = 0
    C1 = np.float32(0.01 ** 2)
    np.testing.assert_allclose(m[0, 15, 15], (2 * mu1 * mu2 + C1) / (mu1 ** 2 + mu2 ** 2 + C1), rtol=1e-4)
    assert m[0, 0, 0] != pytest.approx(m[0, 15, 15], rel=1e-3)                # border sees the zero padding


def test_adam_closed_form(oracle32):
    p, g = np.array([[1.0, 2.0]], np.float32), np.array([[0.5, -0.25]], np.float32)
    z = np.zeros_like(p)
    p1, m1, v1 = oracle32.adam(p, g, z, z, np.array([1], np.uint8), lr=0.1)
    np.testing.assert_allclose(m1, 0.1 * g, rtol=1e-6)
    np.testing.assert_allclose(v1, (1 - np.float32(0.999)) * g * g, rtol=1e-6)   # (1.0f - b2) in fp32, like adam.cu:31
    np.testing.assert_allclose(p1, p - 0.1 * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-15), rtol=1e-4)   # no bias correction
    p2, m2, v2 = oracle32.adam(p, g, z, z, np.array([0], np.uint8), lr=0.1)
    assert np.array_equal(p2, p) and not m2.any() and not v2.any()            # invisible -> untouched


def test_knn_regular_grid(oracle32):
    xs = np.stack(np.meshgrid(np.arange(6.0), np.arange(6.0), np.arange(6.0), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    d = oracle32.knn(xs)
    np.testing.assert_allclose(d, 1.0, rtol=1e-6)                             # three nearest neighbours at distance 1


def test_camera_restatements_agree(oracle32):
    for v in range(8):
        R, t = syn.orbit_pose(v)
        a = syn.make_camera(1920, 1080, 1000.0, 1000.0, 960.0, 540.0, R, t)
        b = oracle32.camera(1920, 1080, 1000.0, 1000.0, 960.0, 540.0, R, t)
        for k in ("view", "proj", "campos", "lims"):
            np.testing.assert_allclose(a[k], b[k], rtol=1e-6, atol=1e-6)
        assert a["tanfovx"] == pytest.approx(0.96, rel=1e-6) and a["tanfovy"] == pytest.approx(0.54, rel=1e-6)
        # every camera of the rig looks at (0,0,10): it projects to the principal point
        p = np.array([0, 0, 10.0, 1.0])
        h = a["proj"].reshape(4, 4).T @ p
        assert abs(h[0] / h[3]) < 1e-5 and abs(h[1] / h[3]) < 1e-5 and h[3] > 0


def test_extend_selection_and_initialisation():
    """extend() (gaussian.cpp:499-638) restated for the GPU version of the next round: per-pixel nearest-point de-duplication
    over ALL points, then the in-image / positive-depth / alpha < 0.99 filter, then the initial parameters."""
    import numpy as np
    from oracle.oracle import Oracle
    o = Oracle(np.float64)
    W, H, fx, fy, cx, cy = 8, 6, 4.0, 4.0, 4.0, 3.0
    R, t = np.eye(3), np.zeros(3)
    T = np.ones((H, W))                       # alpha = 0 everywhere ...
    T[3, 5] = 0.005                           # ... except one saturated pixel (alpha 0.995 >= 0.99)
    pts = np.array([[0.0, 0.0, 2.0],          # 0: pixel (4,3), depth 2
                    [0.05, 0.05, 1.0],        # 1: pixel (4,3) too, nearer -> wins over 0
                    [0.1, 0.1, 1.0],          # 2: pixel (4,3), same depth as 1 but later index -> loses the tie
                    [0.5, 0.0, 2.0],          # 3: pixel (5,3): saturated -> dropped
                    [-10.0, 0.0, 1.0],        # 4: pixel (-36,3): outside -> dropped
                    [1.0, 0.5, 4.0],          # 5: pixel (5,3)? (1*4/4+4, .5*4/4+3) = (5,3) farther than 3 -> loses; 3 itself is dropped later
                    [-0.5, -0.5, 2.0],        # 6: pixel (3,2) kept
                    [0.6, 0.6, -2.0],         # 7: behind the camera: pixel (2,1), still takes part (reference quirk) and is kept
                    [-0.5, -0.5, 2.5]])       # 8: pixel (3,2), farther than 6 -> loses
    rsp = np.array([1, 1, 1, 1, 1, 1, 1, 1, 1.0])
    keep = o.extend_select(pts, rsp, R, t, fx, fy, cx, cy, W, H, T)
    assert keep.tolist() == [1, 6, 7]
    rsp[6] = 0.0                              # depth_in_rsp_frame must be > 0
    assert o.extend_select(pts, rsp, R, t, fx, fy, cx, cy, W, H, T).tolist() == [1, 7]
    # a rigid transform is applied before projecting: shifting the camera by +1 in x moves every pixel by fx/z
    keep2 = o.extend_select(pts + np.array([1.0, 0, 0]), np.ones(9), R, np.array([-1.0, 0, 0]), fx, fy, cx, cy, W, H, T)
    assert keep2.tolist() == [1, 6, 7]
    cols = np.linspace(0, 1, 27).reshape(9, 3)
    init = o.extend_init(keep, pts, cols, np.arange(1, 10.0), 0.7, fx, 6.0)
    np.testing.assert_allclose(init["xyz"], pts[keep])
    np.testing.assert_allclose(init["f_dc"], (cols[keep] - 0.5) / 0.28209479177387814, rtol=1e-14)
    np.testing.assert_allclose(init["log_scale"], np.log(0.7 * np.arange(1, 10.0)[keep] / 5.0)[:, None].repeat(3, 1), rtol=1e-14)
    assert np.array_equal(init["rot"], np.tile([1.0, 0, 0, 0], (3, 1)))
    np.testing.assert_allclose(init["opacity_logit"], np.log(0.1 / 0.9), rtol=1e-14)
    # random cross-check against a dictionary restatement of the reference's loop
    rng = np.random.default_rng(0)
    n = 400
    p = np.c_[rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(0.5, 6, n)]
    Tm = rng.uniform(0, 0.05, (H, W)); Tm[rng.uniform(size=(H, W)) < 0.6] = 1.0
    d = rng.uniform(-0.2, 3, n)
    keep = o.extend_select(p, d, R, t, fx, fy, cx, cy, W, H, Tm)
    best = {}
    for i in range(n):
        key = (int(np.floor(p[i, 0] * fx / p[i, 2] + cx)), int(np.floor(p[i, 1] * fy / p[i, 2] + cy)))
        if key not in best or p[i, 2] < best[key][1]:
            best[key] = (i, p[i, 2])
    want = sorted(i for (x, y), (i, _) in best.items() if 0 <= x < W and 0 <= y < H and d[i] > 0 and 1 - Tm[y, x] < 0.99)
    assert keep.tolist() == want


def test_activation_oracle_matches_the_reference_ops_on_cpu():
    """The reference's getters ARE torch ops (gaussian.cpp:147-175: torch::sigmoid, torch::exp, functional::normalize) and
    their backward is autograd's: run exactly those on the CPU and pin the oracle's restatement to them."""
    import numpy as np
    import pytest
    torch = pytest.importorskip("torch")
    from oracle.oracle import Oracle
    o = Oracle(np.float32)
    rng = np.random.default_rng(8)
    P = 2000
    a = rng.normal(0, 2, P).astype(np.float32)
    b = rng.normal(-4, 0.6, (P, 3)).astype(np.float32)
    c = rng.normal(0, 1, (P, 4)).astype(np.float32)
    w1, w2, w3 = (rng.normal(size=s).astype(np.float32) for s in ((P,), (P, 3), (P, 4)))
    ta, tb, tc = (torch.tensor(x, requires_grad=True) for x in (a, b, c))
    top, tsc, trot = torch.sigmoid(ta), torch.exp(tb), torch.nn.functional.normalize(tc)
    (top * torch.tensor(w1)).sum().backward()
    (tsc * torch.tensor(w2)).sum().backward()
    (trot * torch.tensor(w3)).sum().backward()
    op, sc, rot = o.activations(a, b, c)
    np.testing.assert_allclose(op, top.detach().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(sc, tsc.detach().numpy(), rtol=2e-6)
    np.testing.assert_allclose(rot, trot.detach().numpy(), rtol=2e-6, atol=1e-7)
    g1, g2, g3 = o.activations_backward(op, sc, c, w1, w2, w3)
    np.testing.assert_allclose(g1, ta.grad.numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(g2, tb.grad.numpy(), rtol=1e-5)
    scale = np.abs(tc.grad.numpy()).max()
    assert np.abs(g3 - tc.grad.numpy()).max() <= 1e-5 * scale


def test_ssim_oracle_matches_the_conv2d_formulation():
    """loss_utils.h:52-127 (`ssim`, the evaluation metric) is the textbook SSIM written with grouped conv2d; the fused kernel
    (and its oracle restatement, A.9) must produce the same map.  Run the conv2d formulation with torch on the CPU."""
    import numpy as np
    import pytest
    torch = pytest.importorskip("torch")
    from oracle.oracle import Oracle
    o = Oracle(np.float32)
    rng = np.random.default_rng(2)
    CH, H, W = 3, 37, 53
    a = rng.uniform(0, 1, (CH, H, W)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.1, a.shape), 0, 1).astype(np.float32)
    x = torch.arange(11, dtype=torch.float32) - 5
    g = torch.exp(-(x * x) / (2 * 1.5 * 1.5))
    g = g / g.sum()
    win = (g[:, None] @ g[None, :]).expand(CH, 1, 11, 11).contiguous()
    conv = lambda t: torch.nn.functional.conv2d(t, win, padding=5, groups=CH)
    i1, i2 = torch.tensor(a)[None], torch.tensor(b)[None]
    mu1, mu2 = conv(i1), conv(i2)
    s1, s2, s12 = conv(i1 * i1) - mu1 * mu1, conv(i2 * i2) - mu2 * mu2, conv(i1 * i2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ref_map = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    m, _, _, _ = o.ssim(a[None], b[None], train=False)
    np.testing.assert_allclose(m[0], ref_map[0].numpy(), rtol=0, atol=3e-5)
    assert abs(float(m.mean()) - float(ref_map.mean())) < 2e-6


def test_photometric_loss_oracle_matches_torch_autograd():
    """gaussian.cpp:692-697: loss = (1 - 0.2) * l1_loss + 0.2 * (1 - fused_ssim).  Value and dL/dimage of the oracle against torch
    autograd through the conv2d SSIM and torch::abs(...).mean() on the CPU."""
    import numpy as np
    import pytest
    torch = pytest.importorskip("torch")
    from oracle.oracle import Oracle
    o = Oracle(np.float32)
    rng = np.random.default_rng(4)
    CH, H, W = 3, 40, 56
    a = rng.uniform(0, 1, (CH, H, W)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.15, a.shape), 0, 1).astype(np.float32)
    x = torch.arange(11, dtype=torch.float64) - 5
    g = torch.exp(-(x * x) / (2 * 1.5 * 1.5))
    g = g / g.sum()
    win = (g[:, None] @ g[None, :]).expand(CH, 1, 11, 11).contiguous()
    conv = lambda t: torch.nn.functional.conv2d(t, win, padding=5, groups=CH)
    i1 = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    i2 = torch.tensor(b, dtype=torch.float64)
    u1, u2 = i1[None], i2[None]
    mu1, mu2 = conv(u1), conv(u2)
    s1, s2, s12 = conv(u1 * u1) - mu1 * mu1, conv(u2 * u2) - mu2 * mu2, conv(u1 * u2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()
    loss = 0.8 * (i1 - i2).abs().mean() + 0.2 * (1 - ssim)
    loss.backward()
    L, dl = o.loss(a, b, 0.2)
    assert abs(L - float(loss.detach())) < 2e-6
    ref = i1.grad.numpy()
    assert np.abs(dl - ref).max() <= 2e-4 * np.abs(ref).max()
