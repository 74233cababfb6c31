"""ctypes binding of libneat_b200.so (include/neat_b200.h).  No fallback: if the shared library is
missing or a call fails, an exception is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NEAT_LIB_VARIANT=<tag> (measurement scripts only): an A/B build made by `python -m neat_b200.build --variant <tag> -D...`
_VARIANT = os.environ.get("NEAT_LIB_VARIANT")
LIB_PATH = os.path.join(_HERE, "libneat_b200.%s.so" % _VARIANT if _VARIANT else "libneat_b200.so")


class NeatError(RuntimeError):
    pass


class NetConfig(ctypes.Structure):
    _fields_ = [("sdf_layers", ctypes.c_int), ("sdf_hidden", ctypes.c_int), ("sdf_skip", ctypes.c_int),
                ("multires", ctypes.c_int), ("feat", ctypes.c_int), ("head_layers", ctypes.c_int),
                ("head_hidden", ctypes.c_int), ("multires_view", ctypes.c_int),
                ("sphere_radius", ctypes.c_float), ("sphere_scale", ctypes.c_float)]


_P = ctypes.c_void_p
_I = ctypes.c_int
# name -> (restype, argtypes); every symbol declared in include/neat_b200.h
SIGNATURES = {
    "neat_create": (_I, [ctypes.POINTER(NetConfig), ctypes.POINTER(_P)]),
    "neat_destroy": (None, [_P]),
    "neat_last_error": (ctypes.c_char_p, []),
    "neat_launch_count": (ctypes.c_longlong, []),
    "neat_param_count": (ctypes.c_size_t, [_P]),
    "neat_param_offset": (ctypes.c_long, [_P, _I, _I, _I]),
    "neat_layer_dims": (_I, [_P, _I, _I, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    "neat_weight_norm_forward": (_I, [_P, _P, _I, _P, _P]),
    "neat_weight_norm_backward": (_I, [_P, _P, _I, _P, _I, _P]),
    "neat_set_precision": (_I, [_P, _I]),
    "neat_l3d_candidates": (_I, [_I, _P, _P, _P, _P, _P, _P]),
    "neat_pack_weights": (_I, [_P, _P, _P]),
    "neat_sdf_points": (_I, [_P, _P, _I, _P, _P]),
    "neat_sdf_rays": (_I, [_P, _P, _I, _P, _P, _I, _I, _P, _P]),
    "neat_sdf_grid": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "neat_sampler_workspace_bytes": (ctypes.c_size_t, [_I]),
    "neat_sampler_run": (_I, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P]),
    "neat_sampler_finish": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P]),
    "neat_feat_tiles_bytes": (ctypes.c_size_t, [_I]),
    "neat_sdf_save_bytes": (ctypes.c_size_t, [_P, _I, _I]),
    "neat_sdf_outputs": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "neat_head_save_bytes": (ctypes.c_size_t, [_P, _I]),
    "neat_head_forward": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P]),
    "neat_camera_rays": (_I, [_P, _P, _P, _I, _P, _P, _P]),
    "neat_composite_forward": (_I, [_P, _P]),
    "neat_line_geometry": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "neat_encodels": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "neat_point_line_attraction": (_I, [_P, _I, _I, _I, ctypes.c_float, _P, _P, _P, _P]),
    "neat_mask_compact_workspace_bytes": (ctypes.c_size_t, [ctypes.c_longlong]),
    "neat_mask_compact": (_I, [_P, ctypes.c_longlong, _P, _P, _P, _P]),
    "neat_sample_pixels": (_I, [_P, _P]),
    "neat_pixel_permutation": (_I, [ctypes.c_uint, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, _P]),
    "neat_linear_sum_assignment": (_I, [_P, _I, _I, _P, _P]),
    "neat_junction_match": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "neat_project_points": (_I, [_I, _P, _P, _I, _P, _P, _P, _P]),
    "neat_project_points_backward": (_I, [_I, _P, _P, _I, _P, _P, _P, _P, _P]),
    "neat_junction_terms": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "neat_junction_terms_backward": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "neat_line_vote_workspace_bytes": (ctypes.c_size_t, [_I, _I]),
    "neat_line_vote": (_I, [_P, _P, _P, _I, _P, _I, ctypes.c_float, _P, _P, _P, _P, _P]),
    "neat_line_visibility": (_I, [_P, _I, _P, _P, _I, _P, _I, ctypes.c_float, _P, _P, _P]),
    "neat_line_junction_graph": (_I, [_P, _I, _P, _I, ctypes.c_float, _P, _P, _P, _P, _P]),
    "neat_adam_step": (_I, [_P, _I, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _I,
                           ctypes.c_float, _P]),
    "neat_loss_forward_backward": (_I, [_P, _P]),
    "neat_project_calib_backward": (_I, [_I, _P, _P, _P, _P, _P]),
    "neat_dbscan_workspace_bytes": (ctypes.c_size_t, [_I]),
    "neat_dbscan": (_I, [_P, _I, ctypes.c_float, _P, _P, _P, _P]),
    "neat_composite_backward": (_I, [_P, _P]),
    "neat_head_bwd_save_bytes": (ctypes.c_size_t, [_P, _I]),
    "neat_feat_bar_bytes": (ctypes.c_size_t, [_I]),
    "neat_head_backward": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _I, _P]),
    "neat_sdf_bwd_save_bytes": (ctypes.c_size_t, [_P, _I]),
    "neat_sdf_bwd_scratch_bytes": (ctypes.c_size_t, [_P, _I]),
    "neat_sdf_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "neat_weight_gradients": (_I, [_P, _P, _I, _P, _P]),
    "neat_train_draws": (_I, [_P, _I, ctypes.c_float, ctypes.c_ulonglong, _P, _P, _P, _P, _P, _P, _P]),
    "neat_gemm_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P]),
    "neat_colsum_f32": (_I, [_P, _I, _I, _I, _P, _I, _P]),
    "neat_junction_step": (_I, [_P, _I, _I, _P, _P, _P, ctypes.c_float, ctypes.c_float, _P, _P, _P, _P]),
    "neat_adam_step_device": (_I, [_P, _I, _P, _P, _P]),
}


class AdamTensor(ctypes.Structure):
    _fields_ = [("param", _P), ("grad", _P), ("exp_avg", _P), ("exp_avg_sq", _P), ("numel", ctypes.c_longlong)]


class PixelArgs(ctypes.Structure):
    _fields_ = [("R", _I), ("W", _I), ("first", ctypes.c_longlong), ("masked", _P), ("n_masked", _I), ("perm", _P),
                ("seed", ctypes.c_ulonglong), ("step", ctypes.c_ulonglong), ("rgb_image", _P), ("labels", _P),
                ("att_points", _P), ("lines", _P), ("n_lines", _I), ("uv", _P), ("uv_proj", _P), ("rgb", _P),
                ("lines2d", _P), ("labels_out", _P), ("index_out", _P)]


class CompositeBwdArgs(ctypes.Structure):
    _fields_ = [("R", _I), ("S", _I), ("z", _P), ("sdf", _P), ("weights", _P), ("rgb", _P), ("act", _P),
                ("rgb_values_bar", _P), ("lines3d_bar", _P), ("beta_param", _P), ("beta_min", ctypes.c_float),
                ("rgb_pre_bar", _P), ("lines_bar", _P), ("sdf_bar", _P), ("beta_bar", _P), ("bg_color", _P)]


class WnLayer(ctypes.Structure):
    _fields_ = [("g", _P), ("v", _P), ("b", _P), ("gg", _P), ("gv", _P), ("gb", _P), ("rows", _I), ("cols", _I),
                ("w_off", ctypes.c_long), ("b_off", ctypes.c_long)]


class LossArgs(ctypes.Structure):
    _fields_ = [("R", _I), ("n_eik", _I), ("rgb_values", _P), ("rgb_gt", _P), ("lines2d", _P), ("lines2d_calib", _P),
                ("lines_gt", _P), ("labels", _P), ("K3", _P), ("k_ld", _I), ("grad_theta", _P),
                ("eikonal_weight", ctypes.c_float), ("line_weight", ctypes.c_float), ("scratch", _P), ("out", _P),
                ("g_rgb", _P), ("g_calib", _P), ("g_theta", _P)]


class GradGroup(ctypes.Structure):
    _fields_ = [("M", _I), ("sdf_fwd_save", _P), ("sdf_bwd_save", _P), ("feat_tiles", _P),
                ("head_fwd_save", _P * 2), ("head_bwd_save", _P * 2)]


class Points(ctypes.Structure):
    _fields_ = [("x", _P), ("dirs", _P), ("rays_o", _P), ("rays_d", _P), ("z", _P),
                ("o_stride", _I), ("R", _I), ("S", _I), ("M", _I)]


class CompositeArgs(ctypes.Structure):
    _fields_ = [("R", _I), ("S", _I), ("z", _P), ("sdf", _P), ("rgb", _P), ("lines", _P), ("normals", _P),
                ("rays_o", _P), ("rays_d", _P), ("beta_param", _P), ("beta_min", ctypes.c_float),
                ("weights", _P), ("rgb_values", _P), ("lines3d", _P), ("depth", _P), ("points3d", _P),
                ("normal_map", _P), ("bg_color", _P)]


class SamplerConfig(ctypes.Structure):
    _fields_ = [("n_eval", ctypes.c_int), ("n_final", ctypes.c_int), ("n_extra", ctypes.c_int),
                ("beta_iters", ctypes.c_int), ("max_iters", ctypes.c_int), ("near_", ctypes.c_float),
                ("far_", ctypes.c_float), ("eps", ctypes.c_float), ("beta_min", ctypes.c_float)]

_lib = None


def load():
    """Load the shared library (building it is __graft_entry__.build()'s / neat_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NeatError("%s not found: run `python -m neat_b200.build` (nvcc, sm_100a). "
                        "There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.neat_debug_set_l2_prefetch.restype = _I
    lib.neat_debug_set_l2_prefetch.argtypes = [_I]
    lib.neat_debug_set_flags.restype = _I
    lib.neat_debug_set_flags.argtypes = [ctypes.c_uint]
    lib.neat_debug_set_grid_cap.restype = _I
    lib.neat_debug_set_grid_cap.argtypes = [_P, _I]
    lib.neat_debug_set_wgrad_split.restype = _I
    lib.neat_debug_set_wgrad_split.argtypes = [_P, _I, _I]
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise NeatError("neat_b200 call failed (%d): %s" % (code, load().neat_last_error().decode()))
