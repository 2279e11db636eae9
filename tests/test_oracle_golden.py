"""Pins the CPU oracle against golden vectors produced by the REFERENCE ITSELF (its own CUDA sources
compiled for sm_100a, see tests/golden/make_golden.py).  No GPU needed: inputs are regenerated from seeds."""
import os

import numpy as np
import pytest

from helpers import grad_close, image_close, small_scene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    p = os.path.join(GOLD, name + ".npz")
    if not os.path.isfile(p):
        pytest.skip("golden vector %s not generated yet" % name)
    return np.load(p)


@pytest.mark.parametrize("name", ["scene_deg3", "scene_deg0", "scene_deg1", "scene_rot_deg3", "scene_rot_deg2"])
def test_oracle_matches_reference_golden(oracle32, name):
    z = _load(name)
    P, W, H, seed, deg = (int(z[k]) for k in ("P", "W", "H", "seed", "degree"))
    view = int(z["view"]) if "view" in z.files else 0
    pp = tuple(float(x) for x in z["pp"]) if "pp" in z.files else (0.0, 0.0)
    g, cam = small_scene(P, W, H, seed, deg, view=view, pp=pp)
    f = oracle32.forward(g, cam)
    st = oracle32.state(f)
    # integer contract: bit-exact
    assert (f["R"], f["B"]) == (int(z["R"]), int(z["B"]))
    np.testing.assert_array_equal(f["radii"], z["radii"])
    np.testing.assert_array_equal(st["tiles_touched"], z["tiles_touched"].view(np.uint32))
    vis = f["radii"] > 0
    np.testing.assert_array_equal(st["depth"].view(np.uint32) * vis, z["depth_bits"])
    np.testing.assert_array_equal(st["xy"].view(np.uint32) * vis[:, None], z["xy_bits"])
    np.testing.assert_array_equal(st["conic_opacity"].view(np.uint32) * vis[:, None], z["conic_opacity_bits"])
    np.testing.assert_array_equal(st["point_list"], z["point_list"].view(np.uint32))
    np.testing.assert_array_equal(st["ranges"], z["ranges"].view(np.uint32))
    np.testing.assert_array_equal(st["bucket_offsets"], z["bucket_offsets"].view(np.uint32))
    # floating point: 1e-4 abs (north_star)
    image_close(f["color"], z["color"], name + " colour")
    image_close(f["final_T"], z["final_T"], name + " final_T")
    dL = np.random.default_rng(seed).normal(size=(3, H, W)).astype(np.float32)
    b = oracle32.backward(f, dL)
    for n in ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_ddc", "dL_dsh", "dL_dscales", "dL_drots"):
        if z[n].size:
            grad_close(b[n].reshape(z[n].shape), z[n], name + " " + n, rtol=2e-3)
    oracle32.free(f)


def test_oracle_aux_ops_match_reference_golden(oracle32):
    z = _load("aux_ops")
    m, d1, d2, d3 = oracle32.ssim(z["img1"][0], z["img2"][0])
    np.testing.assert_allclose(m, z["ssim_map"][0], atol=2e-5)
    grad_close(d1, z["dm_dmu1"][0], "dm_dmu1", rtol=1e-4)
    grad_close(d2, z["dm_dsigma1_sq"][0], "dm_dsigma1_sq", rtol=1e-4)
    grad_close(d3, z["dm_dsigma12"][0], "dm_dsigma12", rtol=1e-4)
    gi = oracle32.ssim_backward(z["img1"][0], z["img2"][0], z["dmap"][0], z["dm_dmu1"][0], z["dm_dsigma1_sq"][0], z["dm_dsigma12"][0])
    grad_close(gi, z["dL_dimg1"][0], "dL_dimg1", rtol=1e-4)
    np.testing.assert_allclose(oracle32.knn(z["knn_pts"]), z["knn"], rtol=2e-6)
    p, mm, vv = oracle32.adam(z["adam_p"], z["adam_g"], z["adam_m"], z["adam_v"], z["adam_vis"], 1e-3)
    np.testing.assert_allclose(p, z["adam_p2"], rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(mm, z["adam_m2"], rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(vv, z["adam_v2"], rtol=2e-6, atol=1e-12)
