"""ctypes binding of include/egogen_b200.h. The product path has NO CPU fallback: if the shared
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

import torch

from .build import LIB_PATH

_lib = None

c_f32p = C.c_void_p     # device pointers are passed as raw addresses
c_i32p = C.c_void_p


class EgError(RuntimeError):
    pass


class EgLbsModel(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_verts", "n_joints", "n_shape", "n_pose_basis", "n_faces",
                                         "n_hand_pca", "n_extra", "n_landmarks")] + \
               [(n, C.c_void_p) for n in ("v_template", "shapedirs", "posedirs", "J_regressor", "parents",
                                          "lbs_weights", "hand_comp_l", "hand_comp_r", "pose_mean",
                                          "extra_vids", "faces", "lmk_faces_idx", "lmk_bary")]


class EgMotionDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_dim", "h_dim", "z_dim", "mlp_dim", "reg_h", "reg_blocks",
                                         "reg_recur", "body_dim")]


class EgPolicyDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_dim", "ego_dim", "h_dim", "pe_L", "n_blocks", "z_dim")]


class EgCvaeDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_dim", "h_dim", "z_dim", "mlp_dim")]


class EgRegressorDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_dim", "h_dim", "n_blocks", "n_recur", "body_dim")]


class EgEnvConfig(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("finetuning", C.c_int32), ("pene_terminate_count", C.c_int32),
                ("feet_marker_idx", C.c_int32 * 6), ("reproj_factor", C.c_float), ("goal_thresh", C.c_float)] + \
               [(n, C.c_float) for n in ("w_skate", "w_floor", "w_face", "w_look", "w_success", "w_dist",
                                         "w_vp", "w_pene", "ray_len")] + \
               [("pene_mode", C.c_int32), ("map_res", C.c_int32), ("map_extent", C.c_float), ("pene_thres", C.c_float)]


ENV_BUFFER_FIELDS = ("state", "seed", "R0", "T0", "betas", "dist", "steps", "goal", "ego", "obs_dist",
                     "obs_time", "reward", "terminated", "goal_reached", "reward_terms", "out_markers",
                     "out_params", "out_pelvis")


class EgEnvBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ENV_BUFFER_FIELDS]


# name -> (restype, argtypes); the test-suite checks every prototype in the header is listed here
_I, _L, _F, _P = C.c_int, C.c_int64, C.c_float, C.c_void_p
PROTOTYPES = {
    "eg_version": (_I, []),
    "eg_last_error": (C.c_char_p, []),
    "eg_launch_count": (_L, []),
    "eg_profile_enable": (_I, [_I]),
    "eg_stage_profile_enable": (_I, [_I]),
    "eg_stage_profile_read": (_I, [C.POINTER(C.c_double), _I]),
    "eg_profile_read": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "eg_sdf_sample": (_I, [_P, _I, _I, _I, _P, _P, _P, _L, _P, _P, _P]),
    "eg_sdf_prepare": (_I, [_P, _I, _I, _I, _P]),
    "eg_sdf_release": (_I, [_P]),
    "eg_penetration_count": (_I, [_P, _I, _I, _P, _P, _P]),
    "eg_ego_depth": (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _F, _F, _F, _I, _F, _P, _P, _P]),
    "eg_lbs_create": (_I, [C.POINTER(EgLbsModel), _I, C.POINTER(_P)]),
    "eg_lbs_destroy": (None, [_P]),
    "eg_lbs_set_markers": (_I, [_P, _P, _I]),
    "eg_lbs_max_skin_nnz": (_I, [_P]),
    "eg_lbs_set_mainloop": (_I, [_P, _I]),
    "eg_lbs_markers_backward": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "eg_lbs_markers_backward_rot": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "eg_lbs_rest_pelvis": (_I, [_P, _P, _I, _I, _P, _P]),
    "eg_motion_create": (_I, [C.POINTER(EgMotionDims), _P, _I, _I, C.POINTER(_P)]),
    "eg_motion_destroy": (None, [_P]),
    "eg_motion_refresh": (_I, [_P, _P]),
    "eg_motion_set_fused": (_I, [_P, _I]),
    "eg_motion_sample_prior": (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _P, _P]),
    "eg_vposer_create": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "eg_vposer_destroy": (None, [_P]),
    "eg_vposer_encode": (_I, [_P, _P, _I, _I, _P, _P]),
    "eg_matmul": (_I, [_P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "eg_linear_forward": (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _F, _P, _I, _P, _I, _P]),
    "eg_env_create": (_I, [C.POINTER(EgEnvConfig), _P, _P, _P, _I, C.POINTER(_P)]),
    "eg_env_destroy": (None, [_P]),
    "eg_env_set_config": (_I, [_P, C.POINTER(EgEnvConfig)]),
    "eg_env_set_scene": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _I]),
    "eg_env_set_navmesh": (_I, [_P, _P, _I]),
    "eg_env_set_crowd": (_I, [_P, _P, _I, _P, _I]),
    "eg_env_step": (_I, [_P, C.POINTER(EgEnvBuffers), _P, _I, _P]),
    "eg_env_reset": (_I, [_P, C.POINTER(EgEnvBuffers), _P, _I, _P, _P, _P, _P, _P]),
    "eg_env_reset_masked": (_I, [_P, C.POINTER(EgEnvBuffers), _P, _I, _P, _P, _P, _P, _P]),
    "eg_egosensing": (_I, [_P, _P, _P, _I, _P, _I, C.c_double, _P, _I, _P, _P]),
    "eg_env_restart_from_pool": (_I, [C.POINTER(EgEnvBuffers), C.POINTER(EgEnvBuffers), _I, _P, _I, _P, _P]),
    "eg_policy_param_count": (_L, [C.POINTER(EgPolicyDims), C.POINTER(C.c_int64)]),
    "eg_policy_param_offsets": (_I, [C.POINTER(EgPolicyDims), C.POINTER(C.c_int64), _I]),
    "eg_policy_create": (_I, [C.POINTER(EgPolicyDims), _P, _P, _I, C.POINTER(_P)]),
    "eg_policy_destroy": (None, [_P]),
    "eg_policy_forward": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "eg_gauss_sample": (_I, [_P, _P, _I, _I, _F, _F, _P, _P, _P]),
    "eg_ppo_loss_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _F, _I, _P, _P]),
    "eg_ppo_loss_backward_mlp": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _F, _I, _P, _P]),
    "eg_ppo_backward_encoders": (_I, [_P, _P, _I, _P]),
    "eg_moments": (_I, [_P, _L, _P, _P]),
    "eg_adv_normalize": (_I, [_P, _I, _P, _F, _P, _P]),
    "eg_clip_adamw_step": (_I, [_P, _P, _P, _F, _F, _F, _F, _F, _F, _I, _P]),
    "eg_dp_reduce_norm": (_I, [_P, _P, _I, _I, _L, _L, _P, _P, _P, _P]),
    "eg_dp_reduce_range": (_I, [_P, _P, _I, _I, _L, _L, _L, _L, _I, _P, _P, _P, _P]),
    "eg_dp_adamw_gather": (_I, [_P, _P, _I, _I, _L, _L, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _I, _P]),
    "eg_gae": (_I, [_P, _P, _P, _P, _P, _I, _I, C.c_double, C.c_double, _P, _P, _P]),
    "eg_cvae_param_count": (_L, [C.POINTER(EgCvaeDims)]),
    "eg_cvae_create": (_I, [C.POINTER(EgCvaeDims), _P, _P, _I, C.POINTER(_P)]),
    "eg_cvae_destroy": (None, [_P]),
    "eg_cvae_loss_backward": (_I, [_P, _P, _P, _P, _I, _F, _F, _F, _I, _F, _P, _P, _P]),
    "eg_regressor_param_count": (_L, [C.POINTER(EgRegressorDims)]),
    "eg_regressor_train_create": (_I, [C.POINTER(EgRegressorDims), _P, _P, _P, _I, C.POINTER(_P)]),
    "eg_regressor_train_destroy": (None, [_P]),
    "eg_regressor_loss_backward": (_I, [_P, _P, _P, _I, _F, _P, _P, _P]),
    "eg_cvae_forward_train": (_I, [_P, _P, _P, _P, _I, _P, _P]),
    "eg_cvae_backward": (_I, [_P, _P, _P, _P, _I, _F, _F, _F, _I, _F, _I, _P, _P, _P, _P]),
    "eg_regressor_cycle_backward": (_I, [_P, _P, _P, _P, _I, _I, _F, _F, _F, _F, _I, _P, _P, _P, _P]),
    "eg_adam_step_flat": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P]),
    "eg_update_transl_glorot": (_I, [_P, _P, _P, _P, _I, _P, _I, _P, _P, _P]),
    "eg_new_coordinate": (_I, [_P, _I, _I, _P, _P, _P]),
    "eg_rigid_points": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "eg_lbs_forward": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "eg_lbs_forward_sdf": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
}


def lib():
    """Load libegogen_b200.so (once). Raises EgError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EgError(f"{LIB_PATH} is missing - run `python -m egogen_b200.build` (no CPU fallback exists)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise EgError(f"egogen_b200 error {rc}: {lib().eg_last_error().decode()}")


def ptr(t):
    """Raw device address of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EgError("expected a CUDA tensor - the egogen_b200 operators have no CPU path")
    if not t.is_contiguous():
        raise EgError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def f32c(t, device):
    """float32 contiguous copy/view on `device`."""
    return torch.as_tensor(t, dtype=torch.float32, device=device).contiguous()
