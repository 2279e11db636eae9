"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np

from gaussian_lic_b200 import synthetic as syn

# Tolerances (north_star): RGB / transmittance within 1e-4 abs, tile/key indexing bit-exact.
COLOR_ATOL = 1e-4
# expf differs between glibc and the GPU's ex2.approx path by <= 2 ulp; a pixel whose alpha sits within
# that distance of the 1/255 or T<1e-4 thresholds may take the other branch on the CPU oracle.
# Such flips move a pixel by <= alpha*T*c ~ 4e-3; we allow a vanishing fraction of them, bounded in size.
FLIP_FRACTION = 2e-5
FLIP_MAX = 2e-2
GRAD_RTOL = 2e-3       # relative to the max |grad| of the tensor (fp32 atomics reorder sums)


def small_scene(P=3000, W=320, H=208, seed=11, deg=3, zmax=12.0, f=250.0, log_scale_mean=-3.0, view=0, pp=(0.0, 0.0),
                radius=2.0):
    """Gaussians in the world (= view-0 camera) frame + the camera of rig view `view` (synthetic.orbit_pose: view 0 is
    the identity pose, v >= 1 sits on a circle of `radius` around (0,0,10) and looks at it => R != I, t != 0).
    pp = principal-point offset in pixels from the image centre (cx = W/2 + pp[0], cy = H/2 + pp[1]): a non-zero offset
    makes all four lim* clamps of camera.h:63-66 differ and P02 / P12 of the projection non-zero."""
    g = syn.make_gaussians(P, W, H, f, f, sh_degree=deg, zmax=zmax, seed=seed, log_scale_mean=log_scale_mean)
    R_wc, t_wc = syn.orbit_pose(view, radius=radius)
    cam = syn.make_camera(W, H, f, f, W / 2.0 + pp[0], H / 2.0 + pp[1], R_wc, t_wc)
    return g, cam


# (view, principal-point offset): the identity pose, rotated + translated rig views, and an off-centre principal point
POSES = [(0, (0.0, 0.0)), (1, (0.0, 0.0)), (3, (0.0, 0.0)), (5, (17.5, -9.25)), (0, (-23.0, 11.5))]


def image_close(a, b, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b)
    bad = err > COLOR_ATOL
    frac = bad.mean()
    msg = "%s: max err %.3e, frac>1e-4 = %.3e (%d px)" % (what, err.max(), frac, bad.sum())
    print(msg)
    assert frac <= FLIP_FRACTION, msg
    assert err.max() <= FLIP_MAX, msg


# Element-wise criterion on top of the norm-relative one: |a-b| <= EL_RTOL*|b| + EL_ATOL*max|b| for every entry, so that
# small-magnitude entries (most of dL_dsh, dL_drots) are constrained too.  A gradient entry is a sum of up to thousands of
# fp32 terms of both signs, so its rounding noise scales with the tensor (the EL_ATOL term: the f32 oracle differs from
# the f64 oracle by <= 5e-6*max|b| there); EL_FRAC entries may miss it (a pixel whose alpha sits on the 1/255 or T<1e-4
# threshold takes the other branch and moves the Gaussians it touches by a whole pixel contribution).
EL_RTOL = 5e-3
EL_ATOL = 5e-5
EL_FRAC = 1e-3


def grad_close(a, b, what, rtol=GRAD_RTOL, el_rtol=EL_RTOL, el_atol=EL_ATOL, el_frac=EL_FRAC):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.isfinite(a).all(), what + " has non-finite values"
    scale = max(np.abs(b).max(), 1e-12)
    diff = np.abs(a - b)
    err = diff.max() / scale
    bad = diff > el_rtol * np.abs(b) + el_atol * scale
    frac = float(bad.mean()) if bad.size else 0.0
    big = np.abs(b) > 1e-3 * scale
    rel_big = float((diff[big] / np.abs(b)[big]).max()) if big.any() else 0.0
    print("%s: max|diff|/max|ref| = %.3e (scale %.3e); element-wise: %.3e of %d entries beyond %.0e*|ref| + %.0e*max, "
          "max rel err over entries > 1e-3*max = %.3e" % (what, err, scale, frac, bad.size, el_rtol, el_atol, rel_big))
    assert err <= rtol, "%s rel err %.3e > %.1e" % (what, err, rtol)
    assert frac <= el_frac, "%s: %.3e of the entries fail the element-wise bound (allowed %.1e)" % (what, frac, el_frac)
