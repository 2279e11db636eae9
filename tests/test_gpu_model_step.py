"""Packed model step: fused activations, their in-place chain rule, one-launch masked Adam, and the
native mapping iteration built from them.  Oracles: torch's own sigmoid / exp / normalize (+ autograd) -- which is what
the reference's GaussianModel getters are (gaussian.cpp:147-175) -- and the per-group glic_adam_update kernel, itself
pinned against the reference's adamUpdateCUDA in test_gpu_reference_pin.py."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _raw(P, M, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.randn(*s, device="cuda", generator=gen)
    return dict(op=2.0 * r(P), ls=-4.0 + 0.6 * r(P, 3), rot=r(P, 4), gen=gen)


def test_activations_match_torch_forward_and_backward():
    from gaussian_lic_b200 import capi
    lib, P = capi.lib, 10007
    x = _raw(P, 0)
    x["rot"][5] = 0.0                                        # degenerate quaternion: normalize's eps branch
    op, sc, rot = torch.empty(P, device="cuda"), torch.empty(P, 3, device="cuda"), torch.empty(P, 4, device="cuda")
    capi.check(lib.glic_activations_forward(P, capi.ptr(x["op"]), capi.ptr(x["ls"]), capi.ptr(x["rot"]), capi.ptr(op),
                                            capi.ptr(sc), capi.ptr(rot), None), "act")
    a, b, c = (t.clone().requires_grad_(True) for t in (x["op"], x["ls"], x["rot"]))
    top, tsc, trot = torch.sigmoid(a), torch.exp(b), torch.nn.functional.normalize(c)
    torch.testing.assert_close(op, top.detach(), rtol=2e-6, atol=1e-7)
    torch.testing.assert_close(sc, tsc.detach(), rtol=2e-6, atol=0)
    torch.testing.assert_close(rot, trot.detach(), rtol=2e-6, atol=1e-7)
    g_op, g_sc, g_rot = torch.randn(P, device="cuda"), torch.randn(P, 3, device="cuda"), torch.randn(P, 4, device="cuda")
    (top * g_op).sum().backward(); (tsc * g_sc).sum().backward(); (trot * g_rot).sum().backward()
    capi.check(lib.glic_activations_backward(P, capi.ptr(op), capi.ptr(sc), capi.ptr(x["rot"]), capi.ptr(g_op),
                                             capi.ptr(g_sc), capi.ptr(g_rot), None), "act_bwd")
    torch.testing.assert_close(g_op, a.grad, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(g_sc, b.grad, rtol=1e-5, atol=0)
    keep = torch.ones(P, dtype=torch.bool, device="cuda"); keep[5] = False
    scale = float(c.grad[keep].abs().max())
    assert float((g_rot[keep] - c.grad[keep]).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("P,M", [(5003, 15), (4096, 0)])
def test_packed_adam_is_six_group_adams(P, M):
    from gaussian_lic_b200 import capi
    lib = capi.lib
    n = int(lib.glic_packed_floats(P, M))
    off = (C.c_size_t * 6)()
    capi.check(lib.glic_packed_offsets(P, M, off), "off")
    k = (4, 3, 3, 1, 3, 3 * M)
    assert [int(o) for o in off] == list(np.cumsum([0] + [P * kk for kk in k[:-1]])) and n == P * sum(k)
    gen = torch.Generator(device="cuda").manual_seed(3)
    r = lambda: torch.randn(n, device="cuda", generator=gen)
    param, grad, m, v = r(), r() * 1e-3, r() * 1e-3, (r() * 1e-3) ** 2
    vis = (torch.rand(P, device="cuda", generator=gen) < 0.7).to(torch.uint8)
    lrs = [1e-3, 1.6e-4, 5e-3, 5e-2, 2.5e-3, 1.25e-4]
    p1, m1, v1 = param.clone(), m.clone(), v.clone()
    capi.check(lib.glic_adam_update_packed(capi.ptr(p1), capi.ptr(grad), capi.ptr(m1), capi.ptr(v1), capi.ptr(vis),
                                           (C.c_float * 6)(*lrs), 0.9, 0.999, 1e-15, P, M, None), "packed")
    p2, m2, v2 = param.clone(), m.clone(), v.clone()
    for q in range(6):
        if k[q] == 0:
            continue
        s = slice(int(off[q]), int(off[q]) + P * k[q])
        capi.check(lib.glic_adam_update(capi.ptr(p2[s]), capi.ptr(grad[s]), capi.ptr(m2[s]), capi.ptr(v2[s]), capi.ptr(vis),
                                        lrs[q], 0.9, 0.999, 1e-15, P, k[q], None), "adam")
    torch.cuda.synchronize()
    assert torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)
    untouched = (vis == 0).repeat_interleave(4)
    assert torch.equal(p1[:4 * P][untouched], param[:4 * P][untouched])


def test_native_iteration_descends_and_graph_replays():
    from gaussian_lic_b200 import model, ops, synthetic as syn
    from helpers import small_scene
    P, W, H = 6000, 320, 208
    g, cam = small_scene(P, W, H, 21, 3)
    dev = torch.device("cuda:0")
    mdl = model.PackedModel(g, dev)
    rast = ops.CRasterizer(W, H, dev)
    view = rast.make_view(cam)
    gt = torch.as_tensor(syn.make_gt_image(W, H)).to(dev)
    f32 = dict(dtype=torch.float32, device=dev)
    color, T = torch.empty(3, H, W, **f32), torch.empty(H, W, **f32)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    loss, dL = torch.empty(1, **f32), torch.empty(3, H, W, **f32)
    # first activation must reproduce the scene's own activated parameters
    mdl.activate()
    np.testing.assert_allclose(mdl.act_opacity.cpu().numpy(), np.asarray(g["opacity"], np.float32).reshape(-1), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mdl.act_scales.cpu().numpy(), np.asarray(g["scales"], np.float32), rtol=1e-5, atol=0)
    rast.forward(mdl.inputs, view, out_color=color, out_T=T, radii=radii, sync=True)      # settles the binning capacity
    losses = []
    for _ in range(6):
        mdl.iteration(rast, view, gt, color, T, radii, loss, dL)
        losses.append(float(loss.item()))
    assert not rast.finish()
    # same sequence captured once and replayed
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        rast.stream = side.cuda_stream
        mdl.iteration(rast, view, gt, color, T, radii, loss, dL)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            rast.stream = torch.cuda.current_stream(dev).cuda_stream
            mdl.iteration(rast, view, gt, color, T, radii, loss, dL)
    rast.stream = None
    torch.cuda.current_stream(dev).wait_stream(side)
    for _ in range(12):
        graph.replay()
    torch.cuda.synchronize()
    losses.append(float(loss.item()))
    assert not rast.finish()
    assert all(np.isfinite(losses)) and bool(torch.isfinite(mdl.params).all())
    assert losses[-1] < losses[0], losses


def test_model_step_matches_cpu_oracle(oracle32):
    """CUDA activations / chain rule / packed Adam against the CPU restatement (oracle/glic_oracle.c)."""
    from gaussian_lic_b200 import capi
    lib, P, M = capi.lib, 3001, 15
    rng = np.random.default_rng(4)
    a = rng.normal(0, 2, P).astype(np.float32)
    b = rng.normal(-4, 0.6, (P, 3)).astype(np.float32)
    c = rng.normal(0, 1, (P, 4)).astype(np.float32)
    w1, w2, w3 = (rng.normal(size=s).astype(np.float32) for s in ((P,), (P, 3), (P, 4)))
    o_op, o_sc, o_rot = oracle32.activations(a, b, c)
    o_g = oracle32.activations_backward(o_op, o_sc, c, w1, w2, w3)
    t = lambda x: torch.as_tensor(x).cuda()
    da, db, dc_ = t(a), t(b), t(c)
    op, sc, rot = torch.empty(P, device="cuda"), torch.empty(P, 3, device="cuda"), torch.empty(P, 4, device="cuda")
    capi.check(lib.glic_activations_forward(P, capi.ptr(da), capi.ptr(db), capi.ptr(dc_), capi.ptr(op), capi.ptr(sc),
                                            capi.ptr(rot), None), "act")
    g1, g2, g3 = t(w1), t(w2), t(w3)
    capi.check(lib.glic_activations_backward(P, capi.ptr(op), capi.ptr(sc), capi.ptr(dc_), capi.ptr(g1), capi.ptr(g2),
                                             capi.ptr(g3), None), "act_bwd")
    for mine, ref, what in ((op, o_op, "opacity"), (sc, o_sc, "scale"), (rot, o_rot, "rotation"), (g1, o_g[0], "d opacity"),
                            (g2, o_g[1], "d scale"), (g3, o_g[2], "d rotation")):
        err = np.abs(mine.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-12)
        assert err <= 5e-6, (what, err)             # expf / division rounding only
    n = P * 59
    p, g, m, v = (rng.normal(size=n).astype(np.float32) for _ in range(4))
    g *= 1e-3; m *= 1e-3; v = (v * 1e-3) ** 2
    vis = (rng.uniform(size=P) < 0.7).astype(np.uint8)
    lr6 = [1e-3, 1.6e-4, 5e-3, 5e-2, 2.5e-3, 1.25e-4]
    op_, om_, ov_ = oracle32.adam_packed(p, g, m, v, vis, lr6, M)
    dp, dg, dm, dv, dvis = t(p), t(g), t(m), t(v), t(vis)
    capi.check(lib.glic_adam_update_packed(capi.ptr(dp), capi.ptr(dg), capi.ptr(dm), capi.ptr(dv), capi.ptr(dvis),
                                           (C.c_float * 6)(*lr6), 0.9, 0.999, 1e-15, P, M, None), "packed")
    torch.cuda.synchronize()
    # same tolerance as the per-group Adam test: nvcc contracts b1*m + (1-b1)*g into an fma
    np.testing.assert_allclose(dm.cpu().numpy(), om_, rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(dv.cpu().numpy(), ov_, rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(dp.cpu().numpy(), op_, rtol=1e-5, atol=2e-6)
