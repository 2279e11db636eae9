"""Deterministic synthetic scenes for parity tests and benchmarks.

Pure numpy (host side); the same arrays feed the CUDA path, the CPU oracle and the
reference build, so every implementation sees identical bits.

Camera math restates /root/reference/src/camera.h:38-110 (FoV, view / projection matrices
in the column-major layout the kernels receive, asymmetric lim clamps) in closed form.
"""
import math

import numpy as np

CONFIGS = {
    # name: (P, W, H, fx, fy, cx, cy, sh_degree, zmax)
    "cfg1": (10_000, 640, 480, 400.0, 400.0, 320.0, 240.0, 0, 20.0),
    "cfg2": (500_000, 1920, 1080, 1000.0, 1000.0, 960.0, 540.0, 3, 20.0),
    "cfg3": (2_000_000, 1920, 1080, 1000.0, 1000.0, 960.0, 540.0, 3, 20.0),
    "cfg4": (5_000_000, 1920, 1080, 1000.0, 1000.0, 960.0, 540.0, 3, 60.0),
}


def make_camera(W, H, fx, fy, cx, cy, R_wc=None, t_wc=None):
    """Pinhole camera -> the 16+16+3 floats + scalars handed to the rasterizer.

    view[4c+r] = Rt[r][c] and proj[4c+r] = (P*Rt)[r][c] (column-major).
    """
    R_wc = np.eye(3) if R_wc is None else np.asarray(R_wc, np.float64)
    t_wc = np.zeros(3) if t_wc is None else np.asarray(t_wc, np.float64)
    R_cw = R_wc.T
    t_cw = -R_wc.T @ t_wc
    Rt = np.eye(4, dtype=np.float32)
    Rt[:3, :3] = R_cw.astype(np.float32)
    Rt[:3, 3] = t_cw.astype(np.float32)
    fovx = np.float32(2.0 * math.atan(W / (2.0 * fx)))
    fovy = np.float32(2.0 * math.atan(H / (2.0 * fy)))
    znear, zfar = np.float32(0.01), np.float32(100.0)
    Pm = np.zeros((4, 4), np.float32)
    Pm[0, 0] = np.float32(1.0 / math.tan(float(fovx) / 2))
    Pm[1, 1] = np.float32(1.0 / math.tan(float(fovy) / 2))
    Pm[0, 2] = (np.float32(2) * np.float32(cx) - np.float32(W)) / np.float32(W)
    Pm[1, 2] = (np.float32(2) * np.float32(cy) - np.float32(H)) / np.float32(H)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    full = (Pm @ Rt).astype(np.float32)
    campos = (-(R_cw.T @ t_cw)).astype(np.float32)
    ffx, ffy, fcx, fcy = np.float32(fx), np.float32(fy), np.float32(cx), np.float32(cy)
    lims = np.array([-0.15 * W / ffx - fcx / ffx, 1.15 * W / ffx - fcx / ffx,
                     -0.15 * H / ffy - fcy / ffy, 1.15 * H / ffy - fcy / ffy], np.float32)
    return dict(W=int(W), H=int(H), view=np.ascontiguousarray(Rt.T).reshape(16).copy(),
                proj=np.ascontiguousarray(full.T).reshape(16).copy(), campos=campos,
                tanfovx=float(np.tan(np.float32(fovx * np.float32(0.5)))),
                tanfovy=float(np.tan(np.float32(fovy * np.float32(0.5)))), lims=lims)


def orbit_pose(v, n_views=8, radius=2.0, target=(0.0, 0.0, 10.0)):
    """View v of the closed-form multi-view rig; view 0 is the identity pose."""
    if v == 0:
        return np.eye(3), np.zeros(3)
    ang = 2.0 * math.pi * v / n_views
    tgt = np.asarray(target, np.float64)
    pos = np.array([radius * math.sin(ang), 0.0, radius * (1.0 - math.cos(ang))])
    fwd = tgt - pos
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, -1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R_wc = np.stack([right, down, fwd], axis=1)  # camera axes (x right, y down, z forward) in world
    return R_wc, pos


def make_gaussians(P, W, H, fx, fy, sh_degree=3, zmax=20.0, seed=42, log_scale_mean=-4.0):
    """Activated Gaussian parameters in the view-0 (= world) frame."""
    rng = np.random.Generator(np.random.Philox(seed))
    z = rng.uniform(-1.0, zmax, P)
    x = np.abs(z) * rng.uniform(-1.1 * W / (2 * fx), 1.1 * W / (2 * fx), P)
    y = np.abs(z) * rng.uniform(-1.1 * H / (2 * fy), 1.1 * H / (2 * fy), P)
    means = np.stack([x, y, z], 1).astype(np.float32)
    log_scales = rng.normal(log_scale_mean, 0.6, (P, 3)).astype(np.float32)
    q = rng.normal(0.0, 1.0, (P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    logits = rng.normal(0.0, 2.0, P)
    logits[rng.random(P) < 0.02] = -7.0
    dc = rng.normal(0.0, 1.0, (P, 3)).astype(np.float32)
    M = (sh_degree + 1) ** 2 - 1
    sh = rng.normal(0.0, 0.1, (P, M, 3)).astype(np.float32)
    return dict(means=means, scales=np.exp(log_scales).astype(np.float32), rots=q.astype(np.float32),
                opacity=(1.0 / (1.0 + np.exp(-logits))).astype(np.float32), dc=dc, sh=sh, degree=int(sh_degree),
                log_scales=log_scales, opacity_logits=logits.astype(np.float32))


def make_gt_image(W, H, seed=7):
    """Low-frequency random RGB image in [0,1]: sum of 16 random 2-D cosines per channel."""
    rng = np.random.Generator(np.random.Philox(seed))
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64) / H, np.arange(W, dtype=np.float64) / W, indexing="ij")
    img = np.zeros((3, H, W))
    for ch in range(3):
        for _ in range(16):
            kx, ky = rng.uniform(-6, 6, 2)
            ph = rng.uniform(0, 2 * math.pi)
            img[ch] += rng.uniform(0.2, 1.0) * np.cos(2 * math.pi * (kx * xx + ky * yy) + ph)
        img[ch] = (img[ch] - img[ch].min()) / (img[ch].max() - img[ch].min())
    return img.astype(np.float32)


def make_scene(name="cfg1", P=None, seed=42, view=0, pp=(0.0, 0.0), radius=2.0, **overrides):
    """Gaussians of config `name` + the camera of rig view `view`; pp = principal-point offset (pixels) from the config's."""
    Pn, W, H, fx, fy, cx, cy, deg, zmax = CONFIGS[name]
    P = Pn if P is None else P
    deg = overrides.pop("sh_degree", deg)
    g = make_gaussians(P, W, H, fx, fy, sh_degree=deg, zmax=zmax, seed=seed, **overrides)
    R_wc, t_wc = orbit_pose(view, radius=radius)
    cam = make_camera(W, H, fx, fy, cx + pp[0], cy + pp[1], R_wc, t_wc)
    return g, cam


def make_sort_pairs(n, seed=99, tiles=8160):
    """cfg5: key = (tile << 32) | float_bits(depth in [0.2, 100)), value = index."""
    rng = np.random.Generator(np.random.Philox(seed))
    tile = rng.integers(0, tiles, n, dtype=np.uint64)
    depth = rng.uniform(0.2, 100.0, n).astype(np.float32)
    keys = (tile << np.uint64(32)) | depth.view(np.uint32).astype(np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    return keys, vals
