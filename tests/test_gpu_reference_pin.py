"""Pins against the REFERENCE ITSELF (oracle/_ref = the reference's own .cu files compiled for sm_100a).

1. CUDA path vs reference: same Gaussians/cameras -> radii, R, B, sorted list, ranges bit-exact; colour /
   final_T within 1e-4 abs (north_star); gradients within fp32-atomics tolerance.
2. CPU oracle vs reference: this is what pins oracle/glic_oracle.c (the reference ships no tests / vectors).
Skipped when oracle/_ref/glic_ref_ext.so has not been built (it is built in the authoring container).
"""
import numpy as np
import pytest

from helpers import POSES, grad_close, image_close, small_scene

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ref_forward(ref, g, cam, no_color=False):
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
    P = g["means"].shape[0]
    args = dict(means=t(g["means"]), opac=t(g["opacity"]).view(P, 1), scales=t(g["scales"]), rots=t(g["rots"]),
                dc=t(g["dc"]).view(P, 1, 3), sh=t(g["sh"]), view=t(cam["view"]).view(4, 4), proj=t(cam["proj"]).view(4, 4),
                campos=t(cam["campos"]), bg=torch.zeros(3, device="cuda"), empty=torch.empty(0, device="cuda"))
    lims = [float(x) for x in cam["lims"]]
    out = ref.RasterizeGaussiansCUDA(args["bg"], args["means"], args["empty"], args["opac"], args["scales"], args["rots"], 1.0,
                                     args["empty"], args["view"], args["proj"], cam["tanfovx"], cam["tanfovy"], cam["H"],
                                     cam["W"], lims[0], lims[1], lims[2], lims[3], args["dc"], args["sh"], g["degree"],
                                     args["campos"], False, False, no_color)
    torch.cuda.synchronize()
    return args, out


# P is a multiple of 256 so the reference's tail-thread aliasing race cannot hit Gaussian P-1.
@pytest.mark.parametrize("view,pp", POSES)
@pytest.mark.parametrize("P,W,H,deg,seed", [(4096, 320, 208, 3, 11), (10240, 640, 480, 0, 42), (25600, 800, 450, 3, 8)])
def test_cuda_path_vs_reference(ref_ext, P, W, H, deg, seed, view, pp):
    from gaussian_lic_b200 import ops
    g, cam = small_scene(P, W, H, seed, deg, view=view, pp=pp)
    args, out = _ref_forward(ref_ext, g, cam)
    R, B, color, final_T, radii, geomB, binB, imgB, smpB = out
    r = ops.CRasterizer(W, H)
    gd = ops.scene_to_device(g)
    view = r.make_view(cam)
    c2, T2, rad2 = r.forward(gd, view)
    torch.cuda.synchronize()
    d = r.debug_state()
    assert (r.R, r.B) == (R, B), (r.R, r.B, R, B)
    assert torch.equal(rad2, radii)
    depths, means2D, conic_o, rgb, tiles, offs, clamped, cov3D = ref_ext.slice_geom(geomB, P)
    vis = radii > 0
    assert torch.equal(d["tiles_touched"], tiles)
    assert torch.equal(d["depth"].view(torch.int32)[vis], depths.view(torch.int32)[vis])
    assert torch.equal(d["xy"].view(torch.int32)[vis], means2D.view(torch.int32)[vis])
    assert torch.equal(d["conic_opacity"].view(torch.int32)[vis], conic_o.view(torch.int32)[vis])
    plist, keys = ref_ext.slice_binning(binB, R)
    assert torch.equal(d["point_list"], plist)
    assert torch.equal(d["keys_sorted"], keys)
    ranges, n_contrib, max_contrib, bucket_offsets = ref_ext.slice_image(imgB, H, W)
    assert torch.equal(d["ranges"], ranges)
    assert torch.equal(d["bucket_offsets"], bucket_offsets)
    assert torch.equal(d["n_contrib"], n_contrib)
    assert torch.equal(d["max_contrib"], max_contrib)
    cerr = (c2 - color).abs().max().item()
    terr = (T2 - final_T).abs().max().item()
    print("vs reference: max|dcolor| = %.3e, max|dT| = %.3e, rgb %.3e" % (cerr, terr, (d["rgb"][vis] - rgb[vis]).abs().max().item()))
    assert cerr <= 1e-4 and terr <= 1e-4
    # backward on the same dL_dpix
    gen = torch.Generator(device="cuda").manual_seed(seed)
    dL = torch.randn(3, H, W, device="cuda", generator=gen)
    ref_g = ref_ext.RasterizeGaussiansBackwardCUDA(args["bg"], args["means"], radii, args["empty"], args["scales"], args["rots"],
                                                   1.0, args["empty"], args["view"], args["proj"], cam["tanfovx"], cam["tanfovy"],
                                                   *[float(x) for x in cam["lims"]], dL, args["dc"], args["sh"], g["degree"],
                                                   args["campos"], geomB, R, binB, imgB, B, smpB, 0.0, False)
    mine = r.backward(gd, view, rad2, dL)
    torch.cuda.synchronize()
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_ddc", "dL_dsh", "dL_dscales", "dL_drots"]
    for n, rg in zip(names, ref_g):
        if rg.numel():
            grad_close(mine[n].cpu().numpy().reshape(rg.shape), rg.cpu().numpy(), n + " vs reference", rtol=5e-4)


@pytest.mark.parametrize("view,pp", POSES)
@pytest.mark.parametrize("P,W,H,deg,seed", [(4096, 320, 208, 3, 11), (10240, 640, 480, 0, 42)])
def test_cpu_oracle_vs_reference(ref_ext, oracle32, P, W, H, deg, seed, view, pp):
    """Pins oracle/glic_oracle.c against outputs of the reference itself."""
    g, cam = small_scene(P, W, H, seed, deg, view=view, pp=pp)
    args, out = _ref_forward(ref_ext, g, cam)
    R, B, color, final_T, radii, geomB, binB, imgB, smpB = out
    f = oracle32.forward(g, cam)
    st = oracle32.state(f)
    assert (f["R"], f["B"]) == (R, B)
    np.testing.assert_array_equal(f["radii"], radii.cpu().numpy())
    plist, keys = ref_ext.slice_binning(binB, R)
    np.testing.assert_array_equal(st["point_list"], plist.cpu().numpy().view(np.uint32))
    np.testing.assert_array_equal(st["keys_sorted"], keys.cpu().numpy().view(np.uint64))
    ranges, n_contrib, max_contrib, bucket_offsets = ref_ext.slice_image(imgB, H, W)
    np.testing.assert_array_equal(st["ranges"], ranges.cpu().numpy().view(np.uint32))
    image_close(f["color"], color.cpu().numpy(), "oracle colour vs reference")
    image_close(f["final_T"], final_T.cpu().numpy(), "oracle final_T vs reference")
    gen = torch.Generator(device="cuda").manual_seed(seed)
    dL = torch.randn(3, H, W, device="cuda", generator=gen)
    ref_g = ref_ext.RasterizeGaussiansBackwardCUDA(args["bg"], args["means"], radii, args["empty"], args["scales"], args["rots"],
                                                   1.0, args["empty"], args["view"], args["proj"], cam["tanfovx"], cam["tanfovy"],
                                                   *[float(x) for x in cam["lims"]], dL, args["dc"], args["sh"], g["degree"],
                                                   args["campos"], geomB, R, binB, imgB, B, smpB, 0.0, False)
    b = oracle32.backward(f, dL.cpu().numpy())
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_ddc", "dL_dsh", "dL_dscales", "dL_drots"]
    for n, rg in zip(names, ref_g):
        if rg.numel():
            grad_close(b[n].reshape(rg.shape), rg.cpu().numpy(), "oracle " + n + " vs reference", rtol=2e-3)
    oracle32.free(f)


def test_aux_ops_vs_reference(ref_ext, oracle32):
    from gaussian_lic_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(0)
    a = torch.rand(1, 3, 150, 230, device="cuda", generator=gen)
    b = (a + 0.1 * torch.randn(a.shape, device="cuda", generator=gen)).clamp(0, 1)
    m1 = ops.fusedssim(ops.SSIM_C1, ops.SSIM_C2, a, b, True)
    m2 = ref_ext.fusedssim(ops.SSIM_C1, ops.SSIM_C2, a, b, True)
    for x, y in zip(m1, m2):
        torch.testing.assert_close(x, y, rtol=1e-4, atol=2e-5)
    dmap = torch.randn(a.shape, device="cuda", generator=gen)
    g1 = ops.fusedssim_backward(ops.SSIM_C1, ops.SSIM_C2, a, b, dmap, *m1[1:])
    g2 = ref_ext.fusedssim_backward(ops.SSIM_C1, ops.SSIM_C2, a, b, dmap, *m2[1:])
    torch.testing.assert_close(g1, g2, rtol=1e-4, atol=1e-4)
    pts = torch.randn(20_000, 3, device="cuda", generator=gen) * 3.0
    torch.testing.assert_close(ops.distCUDA2(pts), ref_ext.distCUDA2(pts), rtol=2e-6, atol=0)
    N, M = 4096, 45
    p = torch.randn(N, M, device="cuda", generator=gen); gr = torch.randn(N, M, device="cuda", generator=gen) * 1e-3
    m = torch.zeros(N, M, device="cuda"); v = torch.zeros(N, M, device="cuda")
    vis = torch.rand(N, device="cuda", generator=gen) < 0.6
    pa, ma, va = p.clone(), m.clone(), v.clone()
    pb, mb, vb = p.clone(), m.clone(), v.clone()
    ops.adamUpdate(pa, gr, ma, va, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
    ref_ext.adamUpdate(pb, gr.clone(), mb, vb, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
    torch.testing.assert_close(pa, pb, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(va, vb, rtol=1e-6, atol=1e-15)
