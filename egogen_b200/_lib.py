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


# name -> (restype, argtypes); the test-suite checks every prototype in the header is listed here
_I, _L, _F, _P = C.c_int, C.c_int64, C.c_float, C.c_void_p
PROTOTYPES = {
    "eg_version": (_I, []),
    "eg_last_error": (C.c_char_p, []),
    "eg_launch_count": (_L, []),
    "eg_sdf_sample": (_I, [_P, _I, _I, _I, _P, _P, _P, _L, _P, _P, _P]),
    "eg_penetration_count": (_I, [_P, _I, _I, _P, _P, _P]),
    "eg_ego_depth": (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _F, _F, _F, _I, _F, _P, _P, _P]),
    "eg_lbs_create": (_I, [C.POINTER(EgLbsModel), _I, C.POINTER(_P)]),
    "eg_lbs_destroy": (None, [_P]),
    "eg_lbs_set_markers": (_I, [_P, _P, _I]),
    "eg_lbs_max_skin_nnz": (_I, [_P]),
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
