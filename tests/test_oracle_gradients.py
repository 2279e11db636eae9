"""float64 central finite differences of the CPU oracle's forward vs its analytic backward
(pins the backward maths of A.6-A.8 independently of CUDA)."""
import numpy as np
import pytest

from gaussian_lic_b200 import synthetic as syn


# identity pose; rotated + translated rig views (R != I exercises the view[0..10] terms of cov2d_project and of the
# dJ -> dt -> dmean chain, backward.cu:138-255); off-centre principal point (asymmetric lim* clamps, camera.h:63-66)
@pytest.mark.parametrize("view,pp", [(0, (0.0, 0.0)), (1, (0.0, 0.0)), (3, (0.0, 0.0)), (5, (9.5, -6.25))])
def test_backward_matches_finite_differences(oracle64, view, pp):
    o = oracle64
    W, H = 128, 96
    g = syn.make_gaussians(300, W, H, 100.0, 100.0, sh_degree=3, zmax=8.0, seed=3, log_scale_mean=-2.0)
    g = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in g.items()}
    g["rots"] /= np.linalg.norm(g["rots"], axis=1, keepdims=True)
    g["opacity"] = g["opacity"].reshape(-1, 1)
    R_wc, t_wc = syn.orbit_pose(view, radius=1.5, target=(0.0, 0.0, 6.0))
    cam = syn.make_camera(W, H, 100.0, 100.0, 64.0 + pp[0], 48.0 + pp[1], R_wc, t_wc)
    rng = np.random.default_rng(0)
    wgt = rng.normal(size=(3, H, W))

    def loss(gg):
        f = o.forward(gg, cam)
        L = float((f["color"] * wgt).sum())
        o.free(f)
        return L

    f = o.forward(g, cam)
    b = o.backward(f, wgt)
    vis = np.where(f["radii"] > 0)[0]
    assert len(vis) > 60
    for pn, gn in [("means", "dL_dmeans3D"), ("scales", "dL_dscales"), ("rots", "dL_drots"), ("opacity", "dL_dopacity"),
                   ("dc", "dL_ddc"), ("sh", "dL_dsh")]:
        errs = []
        for _ in range(16):
            i = rng.choice(vis)
            arr = g[pn]
            j = rng.integers(arr[i].size)
            an = b[gn][i].reshape(-1)[j]
            h = 1e-6 * max(1.0, abs(arr[i].reshape(-1)[j]))
            g2 = dict(g)
            a2 = arr.copy(); a2[i].reshape(-1)[j] += h; g2[pn] = a2; Lp = loss(g2)
            a3 = arr.copy(); a3[i].reshape(-1)[j] -= h; g2[pn] = a3; Lm = loss(g2)
            fd = (Lp - Lm) / (2 * h)
            errs.append(abs(fd - an) / max(1e-8, abs(fd), abs(an)))
        errs = np.array(errs)
        # discrete events (tile set / alpha threshold / clamp changes) can break single samples; the bulk must agree
        assert np.median(errs) < 1e-5, (pn, errs)
        assert (errs < 1e-3).mean() >= 0.8, (pn, errs)
    o.free(f)


def test_loss_gradient_finite_differences(oracle64):
    o = oracle64
    rng = np.random.default_rng(1)
    H, W = 24, 30
    gt = rng.random((3, H, W))
    img = np.clip(gt + 0.1 * rng.normal(size=gt.shape), 0.01, 1)
    L, g = o.loss(img, gt)
    for _ in range(20):
        idx = tuple(rng.integers(s) for s in img.shape)
        h = 1e-6
        a = img.copy(); a[idx] += h; Lp, _ = o.loss(a, gt, grad=False)
        a = img.copy(); a[idx] -= h; Lm, _ = o.loss(a, gt, grad=False)
        fd = (Lp - Lm) / (2 * h)
        assert abs(fd - g[idx]) <= 1e-6 * max(1.0, abs(fd) * 1e3), (fd, g[idx])


def test_activation_chain_rule_matches_finite_differences():
    """Oracle of the packed model step: the in-place chain rule of sigmoid / exp / normalize
    (gaussian.cpp:147-175 + autograd) against central differences of the f64 forward."""
    from oracle.oracle import Oracle
    o = Oracle(np.float64)
    rng = np.random.default_rng(12)
    P = 40
    a, b, c = rng.normal(0, 2, P), rng.normal(-4, 0.6, (P, 3)), rng.normal(0, 1, (P, 4))
    w1, w2, w3 = rng.normal(size=P), rng.normal(size=(P, 3)), rng.normal(size=(P, 4))
    op, sc, rot = o.activations(a, b, c)
    np.testing.assert_allclose(op, 1.0 / (1.0 + np.exp(-a)), rtol=1e-14)
    np.testing.assert_allclose(sc, np.exp(b), rtol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(rot, axis=1), 1.0, rtol=1e-14)
    g1, g2, g3 = o.activations_backward(op, sc, c, w1, w2, w3)

    def loss(a_, b_, c_):
        x, y, z = o.activations(a_, b_, c_)
        return (w1 * x).sum() + (w2 * y).sum() + (w3 * z).sum()

    h = 1e-6
    for arr, grad, idxs in ((a, g1, [(i,) for i in range(0, P, 7)]), (b, g2, [(i, i % 3) for i in range(0, P, 7)]),
                            (c, g3, [(i, i % 4) for i in range(0, P, 5)])):
        for idx in idxs:
            plus, minus = arr.copy(), arr.copy()
            plus[idx] += h
            minus[idx] -= h
            args_p = [plus if x is arr else x for x in (a, b, c)]
            args_m = [minus if x is arr else x for x in (a, b, c)]
            fd = (loss(*args_p) - loss(*args_m)) / (2 * h)
            assert abs(fd - grad[idx]) <= 1e-7 * max(1.0, abs(fd)), (idx, fd, grad[idx])
    # the unit quaternion's gradient is tangent: no component along q
    assert np.abs((g3 * rot).sum(1)).max() < 1e-12


def test_packed_adam_oracle_equals_per_group_adam():
    from oracle.oracle import Oracle
    o = Oracle(np.float32)
    rng = np.random.default_rng(3)
    P, M = 37, 15
    n = P * 59
    p, g, m, v = (rng.normal(size=n).astype(np.float32) for _ in range(4))
    v = v * v
    vis = (rng.uniform(size=P) < 0.6).astype(np.uint8)
    lr6 = [1e-3, 1.6e-4, 5e-3, 5e-2, 2.5e-3, 1.25e-4]
    p2, m2, v2 = o.adam_packed(p, g, m, v, vis, lr6, M)
    # rotation block by hand: element j of the block belongs to Gaussian j // 4
    on = np.repeat(vis.astype(bool), 4)
    mm = np.float32(0.9) * m[:4 * P] + (np.float32(1) - np.float32(0.9)) * g[:4 * P]
    vv = np.float32(0.999) * v[:4 * P] + (np.float32(1) - np.float32(0.999)) * g[:4 * P] * g[:4 * P]
    want = p[:4 * P] + np.where(on, -np.float32(lr6[0]) * mm / (np.sqrt(vv) + np.float32(1e-15)), 0).astype(np.float32)
    np.testing.assert_allclose(p2[:4 * P], want, rtol=2e-6, atol=1e-7)
    assert np.array_equal(p2[:4 * P][~on], p[:4 * P][~on]) and np.array_equal(m2[:4 * P][~on], m[:4 * P][~on])
    # sh-rest block starts after 14 floats per Gaussian and has 45 per Gaussian
    on_sh = np.repeat(vis.astype(bool), 45)
    assert np.array_equal(p2[14 * P:][~on_sh], p[14 * P:][~on_sh])
    assert not np.array_equal(p2[14 * P:][on_sh], p[14 * P:][on_sh])
