"""Packed model + native mapping iteration: host-side driver of the C ABI.

The reference's iteration (gaussian.cpp:674-716) is render() -> loss -> backward -> SparseGaussianAdam over six
parameter tensors, with the activations and their backward done by torch ops.  Here the model, its Adam state and its
gradients are three planar buffers in ONE layout (glic_packed_offsets), and an iteration is a fixed sequence of
asynchronous C-ABI launches on one stream:

    glic_activations_forward -> glic_forward -> glic_l1_ssim_loss -> glic_backward (writes the packed gradients)
    -> glic_activations_backward (chain rule in place) -> [glic_p2p_allreduce_mean] -> glic_adam_update_packed

No host synchronisation, no allocation, no torch kernel: the whole iteration replays as one CUDA graph per rank.
PyTorch only owns the memory.  This module is plumbing for tests and bench.py; the product is libglic_b200.so.
"""
import ctypes as C

import torch

from . import capi
from .dist import PackedGrads

GROUPS = ("rotation", "xyz", "scaling", "opacity", "f_dc", "f_rest")          # buffer order
DEFAULT_LRS = dict(xyz=1.6e-4, f_dc=2.5e-3, f_rest=2.5e-3 / 20.0, opacity=0.05, scaling=0.005, rotation=0.001)  # config/fastlivo.yaml:18-22


class PackedModel:
    def __init__(self, g, device, lrs=None, exchange=None, betas=(0.9, 0.999), eps=1e-15):
        """g: raw parameters (means [P,3], log_scales [P,3], rots [P,4] un-normalised, opacity_logits [P], dc [P,3],
        sh [P,M,3], degree).  exchange: a dist.GradAllReduce / P2PGradAllReduce whose buffer receives the gradients."""
        lib = self.lib = capi.lib
        t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device).contiguous()
        self.P = P = int(g["means"].shape[0])
        self.M = M = int(g["sh"].shape[1]) if g["sh"].size else 0
        self.degree = int(g["degree"])
        self.device = device
        n = int(lib.glic_packed_floats(P, M))
        off = (C.c_size_t * 6)()
        capi.check(lib.glic_packed_offsets(P, M, off), "packed_offsets")
        self.offsets = [int(o) for o in off]
        k = (4, 3, 3, 1, 3, 3 * M)
        self.params = torch.empty(n, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        shapes = ((P, 4), (P, 3), (P, 3), (P,), (P, 3), (P, M, 3))
        self.views = {name: self.params[o:o + P * kk].view(*shape) for name, o, kk, shape in zip(GROUPS, self.offsets, k, shapes)}
        self.views["rotation"].copy_(t(g["rots"]))
        self.views["xyz"].copy_(t(g["means"]))
        self.views["scaling"].copy_(t(g["log_scales"]))
        self.views["opacity"].copy_(t(g["opacity_logits"]).view(P))
        self.views["f_dc"].copy_(t(g["dc"]).view(P, 3))
        if M:
            self.views["f_rest"].copy_(t(g["sh"]))
        # activated rasterizer inputs (opacity, scales, unit rotations): written by glic_activations_forward every iteration
        self.act_opacity = torch.empty(P, dtype=torch.float32, device=device)
        self.act_scales = torch.empty(P, 3, dtype=torch.float32, device=device)
        self.act_rots = torch.empty(P, 4, dtype=torch.float32, device=device)
        self.exchange = exchange
        self.packed = exchange.packed if exchange is not None else PackedGrads(P, M, device)
        lrs = dict(DEFAULT_LRS, **(lrs or {}))
        self.lr6 = (C.c_float * 6)(*[lrs[name] for name in GROUPS])
        self.b1, self.b2, self.eps = float(betas[0]), float(betas[1]), float(eps)
        self.inputs = dict(means=self.views["xyz"], scales=self.act_scales, rots=self.act_rots, opacity=self.act_opacity,
                           dc=self.views["f_dc"], sh=self.views["f_rest"], degree=self.degree)

    # ---- the pieces ------------------------------------------------------------------------------------------
    def activate(self, stream=None):
        capi.check(self.lib.glic_activations_forward(self.P, capi.ptr(self.views["opacity"]), capi.ptr(self.views["scaling"]),
                                                     capi.ptr(self.views["rotation"]), capi.ptr(self.act_opacity),
                                                     capi.ptr(self.act_scales), capi.ptr(self.act_rots), stream),
                   "activations_forward")

    def chain_rule(self, stream=None):
        g = self.packed.grads
        capi.check(self.lib.glic_activations_backward(self.P, capi.ptr(self.act_opacity), capi.ptr(self.act_scales),
                                                      capi.ptr(self.views["rotation"]), capi.ptr(g["dL_dopacity"]),
                                                      capi.ptr(g["dL_dscales"]), capi.ptr(g["dL_drots"]), stream),
                   "activations_backward")

    def adam(self, visible, stream=None):
        capi.check(self.lib.glic_adam_update_packed(capi.ptr(self.params), capi.ptr(self.packed.flat), capi.ptr(self.exp_avg),
                                                    capi.ptr(self.exp_avg_sq), capi.ptr(visible), self.lr6, self.b1, self.b2,
                                                    self.eps, self.P, self.M, stream), "adam_update_packed")

    # ---- on-disk map (GaussianModel::saveMap, gaussian.cpp:305-397) ---------------------------------------------------
    def save_map(self, path):
        """point_cloud.ply in the reference's layout, byte-identical to its writer for the same parameters."""
        host = self.params.detach().to("cpu").contiguous()
        capi.check(self.lib.glic_ply_write_packed(str(path).encode(), self.P, self.M, C.c_void_p(host.data_ptr())), "ply_write_packed")

    @staticmethod
    def load_map(path, device, degree=None, **kw):
        """Raw parameters from a Gaussian-LIC / glic map file -> PackedModel (Adam state zeroed)."""
        import numpy as np
        lib = capi.lib
        P, M = C.c_uint32(), C.c_uint32()
        capi.check(lib.glic_ply_read_header(str(path).encode(), C.byref(P), C.byref(M), None), "ply_read_header")
        P, M = P.value, M.value
        a = dict(means=np.zeros((P, 3), np.float32), dc=np.zeros((P, 1, 3), np.float32), sh=np.zeros((P, M, 3), np.float32),
                 opacity_logits=np.zeros(P, np.float32), log_scales=np.zeros((P, 3), np.float32), rots=np.zeros((P, 4), np.float32))
        vp = lambda x: x.ctypes.data_as(C.c_void_p) if x.size else None
        capi.check(lib.glic_ply_read(str(path).encode(), P, M, vp(a["means"]), vp(a["dc"]), vp(a["sh"]), vp(a["opacity_logits"]),
                                     vp(a["log_scales"]), vp(a["rots"])), "ply_read")
        deg = {0: 0, 3: 1, 8: 2, 15: 3}.get(M) if degree is None else degree
        a["dc"] = a["dc"].reshape(P, 3)
        a["degree"] = deg
        return PackedModel(a, device, **kw)

    # ---- one mapping iteration (asynchronous; CUDA-graph capturable) -------------------------------------------
    def iteration(self, rast, view, gt, color, final_T, radii, loss_out, dL_dpix, lambda_dssim=0.2):
        """One iteration on ONE stream: rast.stream when set, else torch's current stream (the C-ABI launches, the visibility
        mask and the exchange must not straddle two streams).  A binning-capacity overflow of an EARLIER iteration (the frame
        rendered empty and Adam stepped on stale momentum) is reported here, with a lag, from the pinned counters: the caller
        regrows (`rast.finish()`) and re-captures its graph.  The native mapper (csrc/mapper.cu) handles this on the device."""
        if int(rast.counters[2]) != 0:
            raise capi.GlicError("binning capacity overflow in an earlier iteration (true R = %d > capacity %d): call rast.finish(), "
                                 "re-capture and redo the step" % (int(rast.counters[0]), rast.cap))
        handle = rast.stream if rast.stream else torch.cuda.current_stream(self.device).cuda_stream
        keep, rast.stream = rast.stream, handle
        try:
            with torch.cuda.stream(torch.cuda.ExternalStream(handle, device=self.device)):
                s = C.c_void_p(handle)
                self.activate(s)
                rast.forward(self.inputs, view, out_color=color, out_T=final_T, radii=radii, sync=False)
                rast.loss(color, gt, lambda_dssim, loss_out, dL_dpix)
                rast.backward(self.inputs, view, radii, dL_dpix, self.packed.grads)
                self.chain_rule(s)
                if self.exchange is not None:
                    _, visible = self.exchange(radii)
                else:
                    visible = self.packed.visible
                    torch.gt(radii[:self.P], 0, out=visible.view(torch.bool))
                self.adam(visible, s)
        finally:
            rast.stream = keep
