"""SMPLXParser - host-side mirror of the reference's motion/models/baseops.py:271-598 with the SMPL-X
forward pass executed by the fused sm_100a LBS kernels behind include/egogen_b200.h.

Same constructor config keys ('n_batch', 'device', 'marker_placement'), same method names,
argument meaning and outputs (torch branch, ``to_numpy=False``; numpy in/out is converted at the
boundary when ``to_numpy=True``). Extra optional config keys: 'body_model_path' (directory holding
smplx/SMPLX_{MALE,FEMALE}.npz) or 'smplx_models' ({'male': arrays, 'female': arrays} in the
egogen_b200.assets.SMPLX_KEYS layout); without either the seeded surrogate model is used.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import _lib, assets


class LbsModel:
    """One SMPL-X gender resident on one GPU (handle from eg_lbs_create)."""

    def __init__(self, arrays: dict, device, marker_vids=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EgError("egogen_b200 LBS runs on CUDA devices only")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        keep = {}
        m = _lib.EgLbsModel()
        a = {k: np.ascontiguousarray(arrays[k]) for k in assets.SMPLX_KEYS}
        m.n_verts, m.n_joints = a["v_template"].shape[0], a["J_regressor"].shape[0]
        m.n_shape, m.n_pose_basis = a["shapedirs"].shape[2], a["posedirs"].shape[0]
        m.n_faces, m.n_hand_pca = a["faces"].shape[0], a["hand_comp_l"].shape[0]
        m.n_extra, m.n_landmarks = a["extra_vids"].shape[0], a["lmk_faces_idx"].shape[0]
        for k in assets.SMPLX_KEYS:
            want = np.int32 if a[k].dtype.kind in "iu" else np.float32
            keep[k] = np.ascontiguousarray(a[k], dtype=want)
            setattr(m, k, keep[k].ctypes.data_as(C.c_void_p))
        self.n_verts, self.n_joints_out = int(m.n_verts), int(m.n_joints + m.n_extra + m.n_landmarks)
        h = C.c_void_p()
        _lib.check(_lib.lib().eg_lbs_create(C.byref(m), idx, C.byref(h)))
        self._h = h
        self.n_markers = 0
        if marker_vids is not None:
            self.set_markers(marker_vids)

    def set_mainloop(self, use_tcgen05: bool):
        """Full-mesh mainloop: tcgen05/TMEM/TMA TF32 tiles (default) or fp32 SIMT tiles."""
        _lib.check(_lib.lib().eg_lbs_set_mainloop(self._h, int(bool(use_tcgen05))))

    def set_markers(self, vids):
        v = np.ascontiguousarray(np.asarray(vids, dtype=np.int32))
        _lib.check(_lib.lib().eg_lbs_set_markers(self._h, v.ctypes.data_as(C.c_void_p), len(v)))
        self.n_markers = len(v)

    def forward(self, xb, betas, want_verts=False, want_joints=True, want_markers=False):
        xb = _lib.f32c(xb, self.device)
        betas = _lib.f32c(betas, self.device).reshape(-1, 10)
        N = xb.shape[0]
        if xb.shape[1] != 93:
            raise _lib.EgError("xb must be [N,93]")
        if betas.shape[0] not in (1, N):
            raise _lib.EgError("betas must have 1 or N rows")
        mk = lambda n: torch.empty(N, n, 3, dtype=torch.float32, device=self.device)
        verts = mk(self.n_verts) if want_verts else None
        joints = mk(self.n_joints_out) if want_joints else None
        markers = mk(self.n_markers) if want_markers else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_lbs_forward(self._h, _lib.ptr(xb), _lib.ptr(betas), betas.shape[0], N,
                                                 _lib.ptr(verts), _lib.ptr(joints), _lib.ptr(markers),
                                                 _lib.stream_ptr(self.device)))
        return verts, joints, markers

    def markers_backward(self, xb, betas, d_markers):
        """Gradient of the marker output w.r.t. the body parameters: d_xb [N,93] for an upstream d_markers [N,M,3]
        (the SMPL-X gradient of the reference's regressor / combo training losses, models_GAMMA_primitive.py:616-631)."""
        xb = _lib.f32c(xb, self.device)
        betas = _lib.f32c(betas, self.device).reshape(-1, 10)
        g = _lib.f32c(d_markers, self.device)
        N = xb.shape[0]
        if xb.shape[1] != 93 or g.shape != (N, self.n_markers, 3) or betas.shape[0] not in (1, N):
            raise _lib.EgError("markers_backward: xb [N,93], betas [1 or N,10], d_markers [N,n_markers,3] expected")
        d_xb = torch.empty(N, 93, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_lbs_markers_backward(self._h, _lib.ptr(xb), _lib.ptr(betas), betas.shape[0], N,
                                                          _lib.ptr(g), _lib.ptr(d_xb), _lib.stream_ptr(self.device)))
        return d_xb

    def forward_sdf(self, xb, betas, frames_per_env, R0, T0, sdf_dict, skip_mask, want_markers=True):
        """Fused LBS -> world transform -> calc_sdf -> feet skip -> per-body count (crowd_env_2f.py:133-177)."""
        from .sdf import _grid3
        xb = _lib.f32c(xb, self.device)
        betas = _lib.f32c(betas, self.device).reshape(-1, 10)
        N = xb.shape[0]
        R0 = _lib.f32c(R0, self.device).reshape(-1, 9)
        T0 = _lib.f32c(T0, self.device).reshape(-1, 3)
        if R0.shape[0] * frames_per_env != N or T0.shape[0] != R0.shape[0]:
            raise _lib.EgError("R0/T0 must have N/frames_per_env rows")
        grid = _grid3(sdf_dict)
        if sdf_dict.get("_prepared_ptr") != grid.data_ptr():     # one-time conservative coarse grid (exact early-out)
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().eg_sdf_prepare(_lib.ptr(grid), grid.shape[0], grid.shape[1], grid.shape[2],
                                                     _lib.stream_ptr(self.device)))
            sdf_dict["_prepared_ptr"] = grid.data_ptr()
            sdf_dict["_prepared_grid"] = grid                     # keep the exact tensor alive
        center = sdf_dict["center"].to(torch.float32).reshape(-1).contiguous()
        scale = sdf_dict["scale"].to(torch.float32).reshape(-1).contiguous()
        counts = torch.empty(N, dtype=torch.int32, device=self.device)
        joints = torch.empty(N, self.n_joints_out, 3, dtype=torch.float32, device=self.device)
        markers = torch.empty(N, self.n_markers, 3, dtype=torch.float32, device=self.device) if want_markers else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_lbs_forward_sdf(
                self._h, _lib.ptr(xb), _lib.ptr(betas), betas.shape[0], N, frames_per_env, _lib.ptr(R0),
                _lib.ptr(T0), _lib.ptr(grid), grid.shape[0], grid.shape[1], grid.shape[2], _lib.ptr(center),
                _lib.ptr(scale), _lib.ptr(skip_mask), _lib.ptr(counts), _lib.ptr(joints), _lib.ptr(markers),
                _lib.stream_ptr(self.device)))
        return counts, joints, markers

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().eg_lbs_destroy(self._h)
                self._h = None
        except Exception:
            pass


_MODEL_CACHE = {}


def get_lbs_model(gender: str, device, body_model_path: Optional[str] = None, arrays: Optional[dict] = None,
                  marker_vids=None) -> LbsModel:
    """The reference builds 6 SMPL-X copies (3 parsers x 2 genders, main_ppo.py:274-293); model buffers
    are read-only, so one resident copy per (gender, device, marker set) is shared instead."""
    dev = torch.device(device)
    key = (gender, str(dev), body_model_path, id(arrays) if arrays is not None else None,
           tuple(marker_vids) if marker_vids is not None else None)
    if key not in _MODEL_CACHE:
        arr = arrays if arrays is not None else assets.get_smplx_model(gender, body_model_path)
        _MODEL_CACHE[key] = LbsModel(arr, dev, marker_vids)
    return _MODEL_CACHE[key]


class SMPLXParser:
    def __init__(self, config):
        for key, val in config.items():
            setattr(self, key, val)
        self.device = torch.device(self.device)
        placement = config.get("marker_placement", "ssm2_67")
        if placement not in ("cmu_41", "ssm2_67"):
            raise NotImplementedError("marker_placement must be cmu_41 or ssm2_67")
        self.marker = assets.marker_ids(placement)
        models = config.get("smplx_models") or {}
        path = config.get("body_model_path")
        self.bm_male = get_lbs_model("male", self.device, path, models.get("male"), self.marker)
        self.bm_female = get_lbs_model("female", self.device, path, models.get("female"), self.marker)

    def _bm(self, gender):
        if gender == "male":
            return self.bm_male
        if gender == "female":
            return self.bm_female
        raise _lib.EgError("gender must be 'male' or 'female'")

    def _t(self, x):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32,
                               device=self.device)

    def forward_smplx(self, betas, gender, xb, to_numpy=True, output_type="markers"):
        bm = self._bm(gender)
        xb_t, betas_t = self._t(xb), self._t(betas)
        if output_type == "markers":
            output = bm.forward(xb_t, betas_t, want_joints=False, want_markers=True)[2]
        elif output_type == "joints":
            output = bm.forward(xb_t, betas_t)[1][:, :22]
        elif output_type == "all_joints":
            output = bm.forward(xb_t, betas_t)[1]
        elif output_type == "vertices":
            output = bm.forward(xb_t, betas_t, want_verts=True, want_joints=False)[0]
        elif output_type == "raw":
            v, j, _ = bm.forward(xb_t, betas_t, want_verts=True, want_joints=True)
            return SimpleNamespace(vertices=v, joints=j)
        else:
            raise NotImplementedError("other output types are not supported")
        if to_numpy:
            output = output.detach().cpu().numpy()
        return output

    def get_jts(self, betas, gender, xb, to_numpy=True):
        return self.forward_smplx(betas, gender, xb, to_numpy, "joints")

    def get_all_jts(self, betas, gender, xb, to_numpy=True):
        return self.forward_smplx(betas, gender, xb, to_numpy, "all_joints")

    def get_markers(self, betas, gender, xb, to_numpy=True):
        return self.forward_smplx(betas, gender, xb, to_numpy, "markers")

    # ---- canonical-frame helpers (baseops.py:465-598) ------------------------------------------
    def _out(self, t, to_numpy):
        return t.detach().cpu().numpy() if to_numpy else t

    def get_new_coordinate(self, betas, gender, xb, to_numpy=True):
        """baseops.py:465-490: canonical frame of each body (x = left->right hip projected on the floor, z up, origin at
        the pelvis). Returns (new_rotmat [b,3,3], new_transl [b,1,3])."""
        jts = self.forward_smplx(betas, gender, xb, False, "joints").contiguous()       # [b,22,3] on the device
        B = jts.shape[0]
        R = torch.empty(B, 3, 3, device=self.device)
        T = torch.empty(B, 3, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_new_coordinate(_lib.ptr(jts), 22 * 3, B, _lib.ptr(R), _lib.ptr(T),
                                                    _lib.stream_ptr(self.device)))
        return self._out(R, to_numpy), self._out(T.view(B, 1, 3), to_numpy)

    def calc_calibrate_offset(self, bm, betas, body_pose, to_numpy=True):
        """baseops.py:494-534: pelvis of the body with zero transl / global_orient, one row per row of body_pose.
        (The pelvis is the root of the kinematic chain, so the result depends on betas only.)"""
        n = body_pose.shape[0]
        be = self._t(betas).reshape(-1, 10).contiguous()
        if be.shape[0] not in (1, n):
            raise _lib.EgError("betas must have 1 or b rows")
        out = torch.empty(n, 3, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_lbs_rest_pelvis(bm._h, _lib.ptr(be), be.shape[0], n, _lib.ptr(out),
                                                     _lib.stream_ptr(self.device)))
        return self._out(out, to_numpy)

    def update_transl_glorot(self, transf_rotmat, transf_transl, betas, gender, xb, to_numpy=True, inplace=True):
        """baseops.py:537-598: the body parameters re-expressed in the frame (transf_rotmat [b,3,3], transf_transl [b,1,3]).
        Device arithmetic follows the reference's torch branch (tgm conversions) for numpy inputs as well."""
        bm = self._bm(gender)
        xb_t = self._t(xb).contiguous()
        n = xb_t.shape[0]
        R = self._t(transf_rotmat).reshape(n, 9).contiguous()
        T = self._t(transf_transl).reshape(n, 3).contiguous()
        be = self._t(betas).reshape(-1, 10).contiguous()
        if be.shape[0] not in (1, n):
            raise _lib.EgError("betas must have 1 or b rows")
        delta = torch.empty(n, 3, device=self.device)
        same = torch.is_tensor(xb) and xb_t.data_ptr() == xb.data_ptr()
        out = xb_t if (inplace and same) else torch.empty_like(xb_t)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_update_transl_glorot(bm._h, _lib.ptr(R), _lib.ptr(T), _lib.ptr(be), be.shape[0],
                                                          _lib.ptr(xb_t), n, _lib.ptr(delta), _lib.ptr(out),
                                                          _lib.stream_ptr(self.device)))
        if to_numpy:
            res = out.detach().cpu().numpy()
            if inplace and isinstance(xb, np.ndarray):
                xb[:] = res
                return xb
            return res
        if inplace and torch.is_tensor(xb) and not same:
            xb.copy_(out)
            return xb
        return out
