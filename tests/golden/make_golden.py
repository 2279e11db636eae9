#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF.

Runs on the GPU box (needs oracle/_ref/glic_ref_ext.so = the reference's own CUDA sources compiled for
sm_100a by oracle/ref_build/build_ref.py).  Inputs are regenerated from seeds by
gaussian_lic_b200/synthetic.py, so only the reference's OUTPUTS are stored.  The CPU suite
(tests/test_oracle_golden.py) then pins oracle/glic_oracle.c against these vectors without a GPU.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy the .npz into tests/golden/
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gaussian_lic_b200 import synthetic as syn  # noqa: E402
from helpers import small_scene  # noqa: E402

SCENES = {  # name: (P, W, H, seed, degree, view, pp) -- P multiple of 256 (reference tail-thread race)
    "scene_deg3": (2048, 160, 112, 21, 3, 0, (0.0, 0.0)),
    "scene_deg0": (1024, 96, 64, 22, 0, 0, (0.0, 0.0)),
    "scene_deg1": (1536, 128, 80, 23, 1, 0, (0.0, 0.0)),
    # rotated + translated rig view with an off-centre principal point (R != I, t != 0, four distinct lim* clamps)
    "scene_rot_deg3": (4096, 160, 112, 24, 3, 3, (7.5, -4.25)),
    "scene_rot_deg2": (3072, 128, 80, 25, 2, 1, (-5.0, 3.5)),
}
ONLY = [a for a in sys.argv[2:]]          # optional: regenerate only these scenes (others keep their committed vectors)


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "glic_ref_ext.so")
    spec = importlib.util.spec_from_file_location("glic_ref_ext", so)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = load_ref()
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
    for name, (P, W, H, seed, deg, view, pp) in SCENES.items():
        if ONLY and name not in ONLY:
            continue
        g, cam = small_scene(P, W, H, seed, deg, view=view, pp=pp)
        lims = [float(x) for x in cam["lims"]]
        a = dict(means=t(g["means"]), opac=t(g["opacity"]).view(P, 1), scales=t(g["scales"]), rots=t(g["rots"]),
                 dc=t(g["dc"]).view(P, 1, 3), sh=t(g["sh"]), view=t(cam["view"]).view(4, 4), proj=t(cam["proj"]).view(4, 4),
                 campos=t(cam["campos"]), bg=torch.zeros(3, device="cuda"), e=torch.empty(0, device="cuda"))
        R, B, color, final_T, radii, geomB, binB, imgB, smpB = ref.RasterizeGaussiansCUDA(
            a["bg"], a["means"], a["e"], a["opac"], a["scales"], a["rots"], 1.0, a["e"], a["view"], a["proj"], cam["tanfovx"],
            cam["tanfovy"], H, W, lims[0], lims[1], lims[2], lims[3], a["dc"], a["sh"], deg, a["campos"], False, False, False)
        depths, means2D, conic_o, rgb, tiles, offs, clamped, cov3D = ref.slice_geom(geomB, P)
        plist, keys = ref.slice_binning(binB, R)
        ranges, n_contrib, max_contrib, bucket_offsets = ref.slice_image(imgB, H, W)
        dL = torch.as_tensor(np.random.default_rng(seed).normal(size=(3, H, W)).astype(np.float32)).cuda()
        grads = ref.RasterizeGaussiansBackwardCUDA(
            a["bg"], a["means"], radii, a["e"], a["scales"], a["rots"], 1.0, a["e"], a["view"], a["proj"], cam["tanfovx"],
            cam["tanfovy"], lims[0], lims[1], lims[2], lims[3], dL, a["dc"], a["sh"], deg, a["campos"], geomB, R, binB, imgB, B,
            smpB, 0.0, False)
        names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_ddc", "dL_dsh", "dL_dscales", "dL_drots"]
        vis = (radii > 0).cpu().numpy()
        c = lambda x: x.detach().cpu().numpy()
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"), P=P, W=W, H=H, seed=seed, degree=deg, view=view, pp=np.asarray(pp, np.float64), R=R, B=B, color=c(color), final_T=c(final_T),
            radii=c(radii), tiles_touched=c(tiles), depth_bits=c(depths).view(np.uint32) * vis, xy_bits=c(means2D).view(np.uint32) * vis[:, None],
            conic_opacity_bits=c(conic_o).view(np.uint32) * vis[:, None], point_list=c(plist), ranges=c(ranges),
            bucket_offsets=c(bucket_offsets), n_contrib=c(n_contrib), dL_dpix_seed=seed,
            **{n: c(x) for n, x in zip(names, grads)})
        print(name, "R", R, "B", B, "visible", int(vis.sum()))
    if ONLY and "aux_ops" not in ONLY:
        return
    # auxiliary operators
    gen = torch.Generator(device="cuda").manual_seed(5)
    img1 = torch.rand(1, 3, 72, 100, device="cuda", generator=gen)
    img2 = (img1 + 0.1 * torch.randn(img1.shape, device="cuda", generator=gen)).clamp(0, 1)
    m, d1, d2, d3 = ref.fusedssim(0.01 ** 2, 0.03 ** 2, img1, img2, True)
    dmap = torch.randn(img1.shape, device="cuda", generator=gen)
    gi = ref.fusedssim_backward(0.01 ** 2, 0.03 ** 2, img1, img2, dmap, d1, d2, d3)
    pts = torch.randn(3000, 3, device="cuda", generator=gen) * torch.tensor([4.0, 1.0, 2.0], device="cuda")
    knn = ref.distCUDA2(pts.contiguous())
    N, M = 512, 45
    p = torch.randn(N, M, device="cuda", generator=gen); gr = torch.randn(N, M, device="cuda", generator=gen) * 1e-2
    mm = torch.randn(N, M, device="cuda", generator=gen) * 1e-3; vv = torch.rand(N, M, device="cuda", generator=gen) * 1e-5
    vis = torch.rand(N, device="cuda", generator=gen) < 0.6
    p2, m2, v2 = p.clone(), mm.clone(), vv.clone()
    ref.adamUpdate(p2, gr.clone(), m2, v2, vis, 1e-3, 0.9, 0.999, 1e-15, N, M)
    c = lambda x: x.detach().cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "aux_ops.npz"), img1=c(img1), img2=c(img2), ssim_map=c(m), dm_dmu1=c(d1),
                        dm_dsigma1_sq=c(d2), dm_dsigma12=c(d3), dmap=c(dmap), dL_dimg1=c(gi), knn_pts=c(pts), knn=c(knn),
                        adam_p=c(p), adam_g=c(gr), adam_m=c(mm), adam_v=c(vv), adam_vis=c(vis), adam_p2=c(p2), adam_m2=c(m2),
                        adam_v2=c(v2))
    print("aux ok")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
