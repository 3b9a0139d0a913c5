"""calc_sdf - same name / arguments / return as the reference's motion/crowd_ppo/utils.py:54-84,
computed by the sm_100a kernel behind eg_sdf_sample (no torch ops on the data path)."""
from __future__ import annotations

import torch

from . import _lib


def _grid3(sdf_dict):
    g = sdf_dict["sdf"]
    g = g.squeeze() if g.dim() != 3 else g
    if g.dim() != 3:
        raise _lib.EgError("sdf grid must be 3-D")
    return g.contiguous()


def calc_sdf(vertices: torch.Tensor, sdf_dict: dict, return_gradient: bool = False, return_index: bool = False):
    """vertices [B,P,3] (world) -> [B,P] signed distance, negative = penetration.
    ``return_index=True`` additionally returns the int32 [B,P,3] base corner indices."""
    if return_gradient:
        raise NotImplementedError("return_gradient is dead code in the reference (utils.py:69-81)")
    if vertices.dim() != 3 or vertices.shape[-1] != 3:
        raise _lib.EgError("vertices must be [B,P,3]")
    dev = vertices.device
    grid = _grid3(sdf_dict)
    pts = vertices.to(torch.float32).contiguous()
    B, P, _ = pts.shape
    out = torch.empty(B, P, dtype=torch.float32, device=dev)
    idx = torch.empty(B, P, 3, dtype=torch.int32, device=dev) if return_index else None
    center = sdf_dict["center"].to(torch.float32).reshape(-1).contiguous()
    scale = sdf_dict["scale"].to(torch.float32).reshape(-1).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().eg_sdf_sample(_lib.ptr(grid), grid.shape[0], grid.shape[1], grid.shape[2],
                                            _lib.ptr(center), _lib.ptr(scale), _lib.ptr(pts), B * P,
                                            _lib.ptr(out), _lib.ptr(idx), _lib.stream_ptr(dev)))
    return (out, idx) if return_index else out


def penetration_count(sdf_values: torch.Tensor, skip_mask: torch.Tensor = None) -> torch.Tensor:
    """crowd_env_2f.py:170-175: per-row count of entries < 0, skipping columns with skip_mask != 0.
    sdf_values [N,V] -> int32 [N]."""
    s = sdf_values.contiguous()
    N, V = s.shape
    out = torch.empty(N, dtype=torch.int32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().eg_penetration_count(_lib.ptr(s), N, V, _lib.ptr(skip_mask), _lib.ptr(out),
                                                   _lib.stream_ptr(s.device)))
    return out


def ego_depth(sdf_dict: dict, cam: torch.Tensor, H: int = 64, W: int = 64, fx: float = 40.0, fy: float = 40.0,
              max_range: float = 7.0, max_steps: int = 64, hit_eps: float = 1e-3, return_steps: bool = False):
    """Config-5 ego-depth sweep: cam [A,12] (eye, right, up, forward) -> depth [A,H,W]."""
    grid = _grid3(sdf_dict)
    cam = cam.to(torch.float32).contiguous()
    A = cam.shape[0]
    depth = torch.empty(A, H, W, dtype=torch.float32, device=cam.device)
    steps = torch.empty(A, H, W, dtype=torch.int32, device=cam.device) if return_steps else None
    center = sdf_dict["center"].to(torch.float32).reshape(-1).contiguous()
    scale = sdf_dict["scale"].to(torch.float32).reshape(-1).contiguous()
    with torch.cuda.device(cam.device):
        _lib.check(_lib.lib().eg_ego_depth(_lib.ptr(grid), grid.shape[0], grid.shape[1], grid.shape[2],
                                           _lib.ptr(center), _lib.ptr(scale), _lib.ptr(cam), A, H, W,
                                           fx, fy, max_range, max_steps, hit_eps, _lib.ptr(depth),
                                           _lib.ptr(steps), _lib.stream_ptr(cam.device)))
    return (depth, steps) if return_steps else depth
