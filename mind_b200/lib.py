"""ctypes binding of libmind_b200.so (C ABI in include/mind_b200.h).

The library is mandatory: there is no CPU or PyTorch fallback.  If it cannot be loaded the
import of mind_b200.predictor fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIND_B200_LIB") or os.path.join(_HERE, "libmind_b200.so")   # override: development builds only

PREC_FP32 = 0
PREC_F16TC = 1


class MindBatch(C.Structure):
    _fields_ = [("n_scenes", C.c_int32),
                ("actor_off", C.POINTER(C.c_int32)),
                ("lane_off", C.POINTER(C.c_int32)),
                ("actors", C.c_void_p),
                ("lanes", C.c_void_p),
                ("rpe", C.POINTER(C.c_void_p)),
                ("ctrs", C.c_void_p),
                ("vecs", C.c_void_p),
                ("tgt_nodes", C.c_void_p),
                ("tgt_rpe", C.c_void_p)]


class MindOutputs(C.Structure):
    _fields_ = [("cls", C.c_void_p), ("reg", C.c_void_p), ("vel", C.c_void_p),
                ("cov_vel", C.c_void_p), ("param", C.c_void_p)]


class MindTreeLevel(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("n_frontier", "n_actor", "obs_len", "pred_len", "ego_idx", "n_tlane")] +
                [("tar_dist_thres", C.c_float)] +
                [(n, C.c_void_p) for n in ("cls", "reg", "vel", "orig", "rot", "ctrs", "vecs", "hpos", "hang", "hvel", "hcov",
                                           "pprob", "cur_t", "tlane", "cpos", "cang", "cvel", "ccov", "gpos", "order", "cprob",
                                           "keep", "tb")])


class MindCostFields(C.Structure):
    _fields_ = [("gx", C.c_int32), ("gy", C.c_int32), ("xs", C.c_void_p), ("ys", C.c_void_p),
                ("n_lane_pts", C.c_int32), ("lane", C.c_void_p), ("n_nodes", C.c_int32), ("n_actor", C.c_int32),
                ("coef_tgt", C.c_void_p), ("mean", C.c_void_p), ("radius", C.c_void_p),
                ("w_ego", C.c_double), ("w_exo", C.c_double), ("exo_cost_offset", C.c_double),
                ("quad", C.c_void_p), ("fields", C.c_void_p)]


class MindIlqrTree(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("parent", C.c_void_p), ("x0", C.c_void_p), ("dt", C.c_double), ("wheelbase", C.c_double),
                ("gx", C.c_int32), ("gy", C.c_int32), ("res", C.c_double), ("field_offset", C.c_void_p),
                ("xs_grid", C.c_void_p), ("ys_grid", C.c_void_p), ("fields", C.c_void_p), ("w_state", C.c_void_p),
                ("des_state", C.c_void_p), ("w_con", C.c_void_p), ("lower", C.c_void_p), ("upper", C.c_void_p),
                ("w_ctrl", C.c_void_p), ("max_iter", C.c_int32), ("us_init", C.c_void_p), ("xs_out", C.c_void_p),
                ("us_out", C.c_void_p), ("iterations", C.c_void_p), ("cost", C.c_void_p)]


class MindTreeUpdate(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("n_new", "n_actor", "n_lane", "n_tlane")] + [("tar_time_ahead", C.c_float)] +
                [(n, C.c_void_p) for n in ("src", "cpos", "cang", "cvel", "ccov", "ttype", "lane_ctrs", "lane_vecs", "tlane",
                                           "tinfo", "npos", "nang", "nvel", "ncov", "norig", "nrot", "nctrs", "nvecs", "actors",
                                           "geom_c", "geom_v", "tgt_nodes", "tgt_rpe", "tgt_pts")])


# every symbol include/mind_b200.h declares (tests check that all of them are exported)
SYMBOLS = ["mind_create", "mind_destroy", "mind_last_error", "mind_build_info", "mind_set_weight",
           "mind_finalize_weights", "mind_set_option", "mind_workspace_bytes", "mind_workspace_bytes_batch", "mind_forward",
           "mind_upload_packed_bytes", "mind_upload_packed", "mind_debug_tap", "mind_launch_count", "mind_graph_replays", "mind_tc_selftest", "mind_debug_fusion_schedule", "mind_debug_edge_init_pack", "mind_debug_conv_fold_pack", "mind_sync_check", "mind_profile_read",
           "mind_tree_level", "mind_tree_update", "mind_tree_last_error", "mind_cost_fields", "mind_cost_fields_last_error", "mind_ilqr_tree_solve", "mind_ilqr_last_error", "mind_debug_field_eval"]

_lib = None


def load(build_if_missing: bool = True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError("libmind_b200.so not built (python -m mind_b200.build)")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    lib.mind_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.mind_create.restype = C.c_int
    lib.mind_destroy.argtypes = [C.c_void_p]
    lib.mind_destroy.restype = None
    lib.mind_last_error.restype = C.c_char_p
    lib.mind_build_info.restype = C.c_char_p
    lib.mind_set_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    lib.mind_set_weight.restype = C.c_int
    lib.mind_finalize_weights.argtypes = [C.c_void_p]
    lib.mind_finalize_weights.restype = C.c_int
    lib.mind_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.mind_set_option.restype = C.c_int
    lib.mind_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.mind_workspace_bytes.restype = C.c_int64
    lib.mind_workspace_bytes_batch.argtypes = [C.c_void_p, C.POINTER(MindBatch)]
    lib.mind_workspace_bytes_batch.restype = C.c_int64
    lib.mind_forward.argtypes = [C.c_void_p, C.POINTER(MindBatch), C.POINTER(MindOutputs), C.c_void_p,
                                 C.c_int64, C.c_void_p]
    lib.mind_forward.restype = C.c_int
    lib.mind_upload_packed_bytes.argtypes = [C.POINTER(C.c_int64), C.c_int32]
    lib.mind_upload_packed_bytes.restype = C.c_int64
    lib.mind_upload_packed.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.c_int64,
                                       C.POINTER(C.c_int64), C.c_void_p]
    lib.mind_upload_packed.restype = C.c_int
    lib.mind_debug_tap.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.mind_debug_tap.restype = C.c_int64
    lib.mind_launch_count.argtypes = [C.c_void_p]
    lib.mind_launch_count.restype = C.c_int64
    lib.mind_graph_replays.argtypes = [C.c_void_p]
    lib.mind_graph_replays.restype = C.c_int64
    lib.mind_tc_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mind_tc_selftest.restype = C.c_int
    lib.mind_debug_fusion_schedule.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                               C.POINTER(C.c_int32)]
    lib.mind_debug_fusion_schedule.restype = C.c_int
    lib.mind_debug_edge_init_pack.argtypes = [C.POINTER(C.c_float)] * 6
    lib.mind_debug_edge_init_pack.restype = C.c_int
    lib.mind_debug_conv_fold_pack.argtypes = [C.POINTER(C.c_float)] + [C.c_int32] * 6 + [C.POINTER(C.c_float), C.c_int64]
    lib.mind_debug_conv_fold_pack.restype = C.c_int
    lib.mind_sync_check.argtypes = [C.c_void_p]
    lib.mind_sync_check.restype = C.c_int
    lib.mind_profile_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.mind_profile_read.restype = C.c_int
    lib.mind_tree_level.argtypes = [C.POINTER(MindTreeLevel), C.c_void_p]
    lib.mind_tree_level.restype = C.c_int
    lib.mind_tree_update.argtypes = [C.POINTER(MindTreeUpdate), C.c_void_p]
    lib.mind_tree_update.restype = C.c_int
    lib.mind_tree_last_error.restype = C.c_char_p
    lib.mind_cost_fields.argtypes = [C.POINTER(MindCostFields), C.c_void_p]
    lib.mind_cost_fields.restype = C.c_int
    lib.mind_cost_fields_last_error.restype = C.c_char_p
    lib.mind_ilqr_tree_solve.argtypes = [C.POINTER(MindIlqrTree)]
    lib.mind_ilqr_tree_solve.restype = C.c_int
    lib.mind_ilqr_last_error.restype = C.c_char_p
    lib.mind_debug_field_eval.argtypes = [C.POINTER(MindIlqrTree), C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_double)]
    lib.mind_debug_field_eval.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError("libmind_b200 %s failed: %s" % (what, load().mind_last_error().decode()))
