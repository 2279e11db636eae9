"""ctypes face of the native mapping host (include/glic_b200.h "Native mapping host", csrc/mapper.cu).

Plumbing for tests and bench.py only: the host of the loop is the C++ object; this class converts numpy arrays to
pointers and keeps the host memory it handed over alive.  The mapper mirrors GaussianModel + extend() + optimize() +
evaluateVisualQuality() + saveMap() of the reference (gaussian.cpp:113-830).
"""
import ctypes as C

import numpy as np

from . import capi

FASTLIVO = dict(position_lr=1.6e-4, feature_lr=2.5e-3, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
                lambda_dssim=0.2, scaling_scale=1.0)                  # config/fastlivo.yaml:15-24


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


def camera_block(W, H, fx, fy, cx, cy, R_wc, t_wc):
    """camera.h:38-110 as the mapper evaluates it -> dict shaped like synthetic.make_camera's."""
    R = (C.c_float * 9)(*[float(x) for x in np.asarray(R_wc, np.float64).reshape(9)])
    t = (C.c_float * 3)(*[float(x) for x in np.asarray(t_wc, np.float64).reshape(3)])
    out = (C.c_float * 41)()
    capi.check(capi.lib.glic_camera_block(int(W), int(H), fx, fy, cx, cy, R, t, out), "camera_block")
    o = np.array(out[:], np.float32)
    return dict(W=int(W), H=int(H), view=o[:16].copy(), proj=o[16:32].copy(), campos=o[32:35].copy(), tanfovx=float(o[35]),
                tanfovy=float(o[36]), lims=o[37:41].copy())


class Mapper:
    def __init__(self, W, H, fx, fy, cx, cy, sh_degree=3, capacity=0, max_iters=100, seed=0, rank=0, world=1, views_per_rank=1,
                 **lrs):
        cfg = capi.MapperConfig()
        cfg.width, cfg.height, cfg.fx, cfg.fy, cfg.cx, cfg.cy = int(W), int(H), fx, fy, cx, cy
        cfg.sh_degree = int(sh_degree)
        for k, v in dict(FASTLIVO, **lrs).items():
            setattr(cfg, k, float(v))
        cfg.capacity, cfg.max_iters, cfg.seed = int(capacity), int(max_iters), int(seed)
        cfg.rank, cfg.world, cfg.views_per_rank = int(rank), int(world), int(views_per_rank)
        self.cfg = cfg
        self.M = (sh_degree + 1) ** 2 - 1
        self.W, self.H = int(W), int(H)
        self._h = C.c_void_p()
        self._keep = []                                   # host arrays the C++ side points into
        capi.check(capi.lib.glic_mapper_create(C.byref(cfg), C.byref(self._h)), "mapper_create")

    def close(self):
        if self._h:
            capi.lib.glic_mapper_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self, g):
        """g: raw parameters as synthetic.make_gaussians returns them (means, dc, sh, opacity_logits, log_scales, rots)."""
        P = int(g["means"].shape[0])
        a = [_f32(g["means"]), _f32(g["dc"]).reshape(P, 3), _f32(g["sh"]), _f32(g["opacity_logits"]), _f32(g["log_scales"]), _f32(g["rots"])]
        capi.check(capi.lib.glic_mapper_initialize(self._h, P, *[_p(x) for x in a]), "mapper_initialize")

    def add_keyframe(self, R_wc, t_wc, image, train=True, pinned=None):
        """image: [3,H,W] float32 numpy array, or a pinned torch tensor (kept alive here)."""
        kf = capi.Keyframe()
        kf.R_wc[:] = [float(x) for x in np.asarray(R_wc, np.float64).reshape(9)]
        kf.t_wc[:] = [float(x) for x in np.asarray(t_wc, np.float64).reshape(3)]
        if hasattr(image, "data_ptr"):
            self._keep.append(image)
            kf.image = image.data_ptr()
        else:
            img = _f32(image)
            self._keep.append(img)
            kf.image = img.ctypes.data
        capi.check(capi.lib.glic_mapper_add_keyframe(self._h, C.byref(kf), int(bool(train))), "mapper_add_keyframe")

    def extend(self, points, colors, depth_rsp):
        pts, col, dep = _f32(points), _f32(colors), _f32(depth_rsp)
        capi.check(capi.lib.glic_mapper_extend(self._h, int(pts.shape[0]), _p(pts), _p(col), _p(dep)), "mapper_extend")
        return self.stats().last_inserted

    def optimize(self, views=None):
        if views is None:
            capi.check(capi.lib.glic_mapper_optimize(self._h, None, 0), "mapper_optimize")
        else:
            arr = (C.c_int * len(views))(*[int(v) for v in views])
            capi.check(capi.lib.glic_mapper_optimize(self._h, arr, len(views)), "mapper_optimize")
        return self.stats()

    def sample_views(self):
        buf = (C.c_int * max(1, self.cfg.max_iters))()
        n = C.c_int()
        capi.check(capi.lib.glic_mapper_sample_views(self._h, buf, C.byref(n)), "mapper_sample_views")
        return list(buf[:n.value])

    def evaluate(self, index, train=True):
        a, b = C.c_float(), C.c_float()
        capi.check(capi.lib.glic_mapper_evaluate(self._h, int(bool(train)), int(index), C.byref(a), C.byref(b)), "mapper_evaluate")
        return a.value, b.value

    def save_map(self, path):
        capi.check(capi.lib.glic_mapper_save_map(self._h, str(path).encode()), "mapper_save_map")

    def stats(self):
        st = capi.MapperStats()
        capi.check(capi.lib.glic_mapper_stats_get(self._h, C.byref(st)), "mapper_stats")
        return st

    def set_optimizer(self, on):
        capi.check(capi.lib.glic_mapper_set_option(self._h, 1, int(bool(on))), "mapper_set_option")

    def set_binning_pairs(self, pairs):
        capi.check(capi.lib.glic_mapper_set_option(self._h, 2, int(pairs)), "mapper_set_option")

    def synchronize(self):
        capi.check(capi.lib.glic_mapper_synchronize(self._h), "mapper_synchronize")

    def download(self, moments=False):
        P, M = int(self.stats().num_gaussians), self.M
        out = dict(means=np.zeros((P, 3), np.float32), dc=np.zeros((P, 3), np.float32), sh=np.zeros((P, M, 3), np.float32),
                   opacity_logits=np.zeros(P, np.float32), log_scales=np.zeros((P, 3), np.float32), rots=np.zeros((P, 4), np.float32))
        m1 = np.zeros(P * (14 + 3 * M), np.float32) if moments else None
        m2 = np.zeros(P * (14 + 3 * M), np.float32) if moments else None
        capi.check(capi.lib.glic_mapper_download(self._h, _p(out["means"]), _p(out["dc"]), _p(out["sh"]), _p(out["opacity_logits"]),
                                                 _p(out["log_scales"]), _p(out["rots"]), _p(m1), _p(m2)), "mapper_download")
        if moments:
            out["exp_avg"], out["exp_avg_sq"] = m1, m2
        return out

    # ---- multi-GPU wiring: torch.distributed (or anything else) only carries the 64-byte handles --------------------------
    def export_handle(self):
        h = C.create_string_buffer(64)
        capi.check(capi.lib.glic_mapper_export(self._h, h), "mapper_export")
        return h.raw

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.cfg.world
        capi.check(capi.lib.glic_mapper_connect(self._h, blob), "mapper_connect")
