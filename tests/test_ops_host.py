"""Host-side glue of the reference-shaped API (no GPU): render() builds the settings, applies the activations and maps the
rasterizer's outputs exactly like rasterizer/renderer.cpp:21-88.  The boundary call is stubbed; everything else is real."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_render_glue_matches_renderer_cpp(monkeypatch):
    from gaussian_lic_b200 import ops, synthetic as syn
    P, W, H = 12, 64, 48
    g = syn.make_gaussians(P, W, H, 50.0, 50.0, sh_degree=3)
    camn = syn.make_camera(W, H, 50.0, 50.0, 32.0, 24.0)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)
    cam = dict(W=W, H=H, tanfovx=camn["tanfovx"], tanfovy=camn["tanfovy"], lims=camn["lims"], view=t(camn["view"]),
               proj=t(camn["proj"]), campos=t(camn["campos"]))
    pc = ops.GaussianParams(t(g["means"]).requires_grad_(True), t(g["dc"]).view(P, 1, 3).requires_grad_(True),
                            t(g["sh"]).requires_grad_(True), t(g["opacity_logits"]).view(P, 1).requires_grad_(True),
                            t(g["log_scales"]).requires_grad_(True), t(g["rots"]).requires_grad_(True), 3, lambda_erank=0.25)
    seen = {}

    def fake_forward(*args):
        (bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D, view, proj, tfx, tfy, h, w, lxn, lxp, lyn, lyp,
         dc, sh, degree, campos, prefiltered, debug, no_color) = args
        seen.update(locals())
        radii = torch.tensor([3, 0] * (P // 2), dtype=torch.int32)
        u8 = torch.zeros(1, dtype=torch.uint8)
        return 7, 2, opacity.sum() * torch.ones(3, h, w), torch.zeros(h, w), radii, u8, u8, u8, u8

    def fake_backward(*args):
        seen["backward_R"], seen["lambda_erank"] = args[22], args[27]
        z = lambda *s: torch.zeros(*s)
        return z(P, 3), z(P, 3), torch.ones(P, 1), z(P, 3), z(P, 6), z(P, 1, 3), z(P, 15, 3), z(P, 3), z(P, 4)

    monkeypatch.setattr(ops, "RasterizeGaussiansCUDA", fake_forward)
    monkeypatch.setattr(ops, "RasterizeGaussiansBackwardCUDA", fake_backward)
    bg = torch.zeros(3)
    image, final_T, screenspace, visible, radii = ops.render(cam, pc, bg, scaling_modifier=1.5)
    # activated inputs reach the boundary; raw ones do not
    np.testing.assert_allclose(seen["opacity"].detach().numpy().ravel(), g["opacity"], rtol=1e-6)
    np.testing.assert_allclose(seen["scales"].detach().numpy(), g["scales"], rtol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(seen["rotations"].detach().numpy(), axis=1), 1.0, rtol=1e-6)
    assert seen["colors"].numel() == 0 and seen["cov3D"].numel() == 0          # empty precomp tensors (rasterizer.cpp:200-201)
    assert (seen["h"], seen["w"], seen["degree"]) == (H, W, 3) and seen["scale_modifier"] == 1.5
    assert seen["view"].shape == (4, 4) and seen["proj"].shape == (4, 4)
    assert abs(seen["tfx"] - camn["tanfovx"]) < 1e-7 and [seen["lxn"], seen["lxp"], seen["lyn"], seen["lyp"]] == [float(x) for x in camn["lims"]]
    assert not seen["prefiltered"] and not seen["debug"] and not seen["no_color"]
    # outputs: (image, final_T, screenspace_points, radii > 0, radii)
    assert image.shape == (3, H, W) and final_T.shape == (H, W) and screenspace.shape == (P, 3) and screenspace.requires_grad
    assert torch.equal(visible, radii > 0) and int(visible.sum()) == P // 2
    # autograd: the op's gradient w.r.t. the activated opacity flows through the sigmoid to the raw logit
    image.sum().backward()
    assert seen["backward_R"] == 7 and seen["lambda_erank"] == 0.25
    s = torch.sigmoid(pc.opacity_.detach())
    torch.testing.assert_close(pc.opacity_.grad, s * (1 - s))                 # fake backward returned dL/dopacity = 1
    assert pc.xyz_.grad is not None and float(pc.xyz_.grad.abs().max()) == 0.0
