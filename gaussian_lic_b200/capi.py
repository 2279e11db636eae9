"""ctypes bindings of the C ABI (include/glic_b200.h) -- plumbing only.

Device memory comes from torch tensors (`tensor.data_ptr()`); every byte of compute happens in
libglic_b200.so.  There is no fallback: if the library is missing, import fails loudly.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libglic_b200.so")

GLIC_OK = 0


class GlicError(RuntimeError):
    pass


class View(C.Structure):
    _fields_ = [("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("limx_neg", C.c_float), ("limx_pos", C.c_float), ("limy_neg", C.c_float), ("limy_pos", C.c_float),
                ("width", C.c_int), ("height", C.c_int)]


class MapperConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("sh_degree", C.c_int), ("position_lr", C.c_float), ("feature_lr", C.c_float), ("opacity_lr", C.c_float),
                ("scaling_lr", C.c_float), ("rotation_lr", C.c_float), ("lambda_dssim", C.c_float), ("scaling_scale", C.c_float),
                ("capacity", C.c_uint32), ("max_iters", C.c_int), ("seed", C.c_uint64), ("rank", C.c_int), ("world", C.c_int),
                ("views_per_rank", C.c_int)]


class Keyframe(C.Structure):
    _fields_ = [("R_wc", C.c_float * 9), ("t_wc", C.c_float * 3), ("image", C.c_void_p)]


class MapperStats(C.Structure):
    _fields_ = [("num_gaussians", C.c_uint32), ("capacity", C.c_uint32), ("iterations", C.c_uint64), ("last_inserted", C.c_uint32),
                ("last_loss", C.c_float), ("mean_visible", C.c_double), ("overflow_regrows", C.c_uint32), ("ms_extend", C.c_float),
                ("ms_optimize", C.c_float)]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "gaussian_lic_b200: %s is missing. Build it with `python gaussian_lic_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, sz, i64, i32, f32, u32 = C.c_void_p, C.c_size_t, C.c_int64, C.c_int, C.c_float, C.c_uint32
    sig = {
        "glic_last_error": (C.c_char_p, []),
        "glic_abi_version": (i32, []),
        "glic_launch_count": (C.c_uint64, []),
        "glic_profile_enable": (i32, [i32]),
        "glic_profile_read": (i32, [C.POINTER(C.c_float), C.POINTER(i32)]),
        "glic_geom_bytes": (sz, [i32]),
        "glic_image_bytes": (sz, [i32, i32]),
        "glic_binning_bytes": (sz, [i64]),
        "glic_sample_bytes": (sz, [i64, i32, i32]),
        "glic_max_buckets": (i64, [i64, i32, i32]),
        "glic_sort_temp_bytes": (sz, [i64]),
        "glic_loss_scratch_bytes": (sz, [i32, i32, i32]),
        "glic_knn_temp_bytes": (sz, [i32]),
        "glic_forward_preprocess": (i32, [i32, i32, i32, vp, vp, f32, vp, vp, vp, vp, C.POINTER(View), i32, vp, vp, sz,
                                          vp, sz, C.POINTER(i64), vp]),
        "glic_binning_capacity": (i64, [sz, sz, i32, i32, i32]),
        "glic_forward": (i32, [i32, i32, i32, vp, vp, f32, vp, vp, vp, vp, C.POINTER(View), i32, vp, vp, sz, vp, sz, vp, sz, vp, sz, vp, vp,
                               vp, vp]),
        "glic_forward_render": (i32, [i32, C.POINTER(View), i32, i64, vp, vp, vp, sz, vp, sz, vp, vp, C.POINTER(i64), vp]),
        "glic_backward": (i32, [i32, i32, i32, vp, vp, f32, vp, vp, vp, C.POINTER(View), vp, i64, vp, vp, vp, vp, vp, f32,
                                vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "glic_sort_pairs_u64_u32": (i32, [i64, i32, vp, vp, vp, vp, vp, sz, vp]),
        "glic_adam_update": (i32, [vp, vp, vp, vp, vp, f32, f32, f32, f32, u32, u32, vp]),
        "glic_fused_ssim": (i32, [i32, i32, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
        "glic_fused_ssim_backward": (i32, [i32, i32, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "glic_l1_ssim_loss": (i32, [i32, i32, i32, f32, vp, vp, vp, vp, vp, sz, vp]),
        "glic_knn_mean_dist2": (i32, [i32, vp, vp, vp, sz, vp]),
        "glic_extend_bytes": (sz, [i32, i32, i32]),
        "glic_extend": (i32, [i32, vp, vp, vp, C.POINTER(f32), C.POINTER(f32), f32, f32, f32, f32, i32, i32, vp, f32, vp, sz, vp, vp,
                              vp, vp, vp, vp, C.POINTER(i32), vp]),
        "glic_ply_bytes": (sz, [u32, u32]),
        "glic_ply_write": (i32, [C.c_char_p, u32, u32, vp, vp, vp, vp, vp, vp]),
        "glic_ply_write_packed": (i32, [C.c_char_p, u32, u32, vp]),
        "glic_ply_read_header": (i32, [C.c_char_p, C.POINTER(u32), C.POINTER(u32), C.POINTER(sz)]),
        "glic_ply_read": (i32, [C.c_char_p, u32, u32, vp, vp, vp, vp, vp, vp]),
        "glic_packed_floats": (sz, [u32, u32]),
        "glic_packed_offsets": (i32, [u32, u32, C.POINTER(sz)]),
        "glic_activations_forward": (i32, [i32, vp, vp, vp, vp, vp, vp, vp]),
        "glic_activations_backward": (i32, [i32, vp, vp, vp, vp, vp, vp, vp]),
        "glic_adam_update_packed": (i32, [vp, vp, vp, vp, vp, C.POINTER(f32), f32, f32, f32, u32, u32, vp]),
        "glic_p2p_buffer_bytes": (sz, [sz, sz]),
        "glic_p2p_model_bytes": (sz, [sz, sz]),
        "glic_p2p_reduce_adam": (i32, [i32, i32, C.POINTER(vp), u32, u32, vp, vp, C.POINTER(f32), f32, f32, f32, vp]),
        "glic_p2p_slice": (i32, [i32, i32, sz, sz, C.POINTER(sz)]),
        "glic_p2p_alloc": (i32, [sz, C.POINTER(vp), C.c_char_p]),
        "glic_p2p_open": (i32, [C.c_char_p, C.POINTER(vp)]),
        "glic_p2p_close": (i32, [vp]),
        "glic_p2p_free": (i32, [vp]),
        "glic_p2p_allreduce_mean": (i32, [i32, i32, C.POINTER(vp), sz, sz, vp]),
        "glic_p2p_check": (i32, [vp, sz, sz, vp]),
        "glic_eval_scratch_bytes": (sz, [i32, i32, i32]),
        "glic_eval_psnr_ssim": (i32, [i32, i32, i32, vp, vp, vp, vp, sz, vp]),
        "glic_arena_append": (i32, [vp, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp, vp, vp]),
        "glic_arena_regrow": (i32, [vp, u32, vp, u32, u32, u32, vp]),
        "glic_camera_block": (i32, [i32, i32, f32, f32, f32, f32, C.POINTER(f32), C.POINTER(f32), C.POINTER(f32)]),
        "glic_mapper_create": (i32, [C.POINTER(MapperConfig), C.POINTER(vp)]),
        "glic_mapper_destroy": (i32, [vp]),
        "glic_mapper_initialize": (i32, [vp, u32, vp, vp, vp, vp, vp, vp]),
        "glic_mapper_add_keyframe": (i32, [vp, C.POINTER(Keyframe), i32]),
        "glic_mapper_extend": (i32, [vp, i32, vp, vp, vp]),
        "glic_mapper_optimize": (i32, [vp, C.POINTER(i32), i32]),
        "glic_mapper_sample_views": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "glic_mapper_evaluate": (i32, [vp, i32, i32, C.POINTER(f32), C.POINTER(f32)]),
        "glic_mapper_save_map": (i32, [vp, C.c_char_p]),
        "glic_mapper_stats_get": (i32, [vp, C.POINTER(MapperStats)]),
        "glic_mapper_download": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "glic_mapper_export": (i32, [vp, C.c_char_p]),
        "glic_mapper_connect": (i32, [vp, C.c_char_p]),
        "glic_mapper_synchronize": (i32, [vp]),
        "glic_mapper_set_option": (i32, [vp, i32, i32]),
        "glic_debug_geom": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "glic_debug_binning": (i32, [i32, vp, i64, i64, vp, vp, vp, vp]),
        "glic_debug_image": (i32, [i32, i32, vp, vp, vp, vp, vp, C.POINTER(i64), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, SYMBOLS = _load()


def check(status, what):
    if status != GLIC_OK:
        raise GlicError("glic_b200 %s failed (%d): %s" % (what, status, lib.glic_last_error().decode()))


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def launch_count():
    return int(lib.glic_launch_count())


STAGES = ["preprocess", "emit", "sort", "ranges", "render_fwd", "loss_fwd", "loss_bwd", "render_bwd", "preprocess_bwd",
          "adam", "zero", "allreduce"]


def profile_enable(on=True):
    check(lib.glic_profile_enable(int(on)), "profile_enable")


def profile_read():
    """-> {stage: (total_ms, count)} since the last profile_enable(True)."""
    ms = (C.c_float * len(STAGES))()
    cnt = (C.c_int * len(STAGES))()
    check(lib.glic_profile_read(ms, cnt), "profile_read")
    return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(STAGES)}
