"""float64 central finite differences of the CPU oracle's forward vs its analytic backward
(pins the backward maths of A.6-A.8 independently of CUDA)."""
import numpy as np

from gaussian_lic_b200 import synthetic as syn


def test_backward_matches_finite_differences(oracle64):
    o = oracle64
    W, H = 128, 96
    g = syn.make_gaussians(300, W, H, 100.0, 100.0, sh_degree=3, zmax=8.0, seed=3, log_scale_mean=-2.0)
    g = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in g.items()}
    g["rots"] /= np.linalg.norm(g["rots"], axis=1, keepdims=True)
    g["opacity"] = g["opacity"].reshape(-1, 1)
    cam = syn.make_camera(W, H, 100.0, 100.0, 64.0, 48.0)
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
    assert len(vis) > 100
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
