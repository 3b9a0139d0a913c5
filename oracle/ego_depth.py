"""Ego-depth sphere-tracing oracle (TEST INFRASTRUCTURE). BASELINE config 5 has NO reference implementation
(the reference's ego-perception is 32 2-D rays; its only depth images are pyrender
rasterisations in experiments/gen_egobody_depth.py) => parity unpinned; this oracle DEFINES the operator:
pinhole rays from the head camera (camera convention of experiments/gen_egobody_depth.py:163-199: eye at the mean of
joints 23/24, gaze from joints 56/57), marched through the reference's calc_sdf field (motion/crowd_ppo/utils.py:54-84):
t <- t + max(d, eps) until d < eps, t > max_range or max_steps."""
import torch

from .sdf import calc_sdf


def camera_from_joints(joints_w: torch.Tensor) -> torch.Tensor:
    """joints_w [A,127,3] world -> cam [A,12] = (eye, right, up, forward); z-up world."""
    eye = (joints_w[:, 23] + joints_w[:, 24]) / 2
    fwd = (joints_w[:, 56] - joints_w[:, 24]) + (joints_w[:, 57] - joints_w[:, 23])
    fwd = fwd / fwd.norm(dim=-1, keepdim=True).clip(min=1e-12)
    up0 = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.cross(fwd, up0, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True).clip(min=1e-12)
    up = torch.cross(right, fwd, dim=-1)
    return torch.cat([eye, right, up, fwd], dim=-1)


def ego_depth(sdf_dict, cam, H=64, W=64, fx=40.0, fy=40.0, max_range=7.0, max_steps=64, hit_eps=1e-3):
    """cam [A,12] -> (depth [A,H,W] float32, steps [A,H,W] int32); float32 arithmetic like the kernel."""
    A = cam.shape[0]
    px = torch.arange(W, dtype=torch.float32)
    py = torch.arange(H, dtype=torch.float32)
    u = ((px + 0.5 - 0.5 * W) / fx).view(1, 1, W)
    v = ((py + 0.5 - 0.5 * H) / fy).view(1, H, 1)
    eye, right, up, fwd = [cam[:, i:i + 3].view(A, 1, 1, 3) for i in (0, 3, 6, 9)]
    d = fwd + u.unsqueeze(-1) * right - v.unsqueeze(-1) * up
    d = d / d.norm(dim=-1, keepdim=True)
    t = torch.zeros(A, H, W)
    steps = torch.zeros(A, H, W, dtype=torch.int32)
    active = torch.ones(A, H, W, dtype=torch.bool)
    for it in range(max_steps):
        if not active.any():
            break
        p = eye + t.unsqueeze(-1) * d
        dist = calc_sdf(p.reshape(1, -1, 3), sdf_dict).reshape(A, H, W)
        hit = active & (dist < hit_eps)
        active = active & ~hit
        t = torch.where(active, t + torch.clamp(dist, min=hit_eps), t)
        out = active & (t > max_range)
        t = torch.where(out, torch.full_like(t, max_range), t)
        active = active & ~out
        steps = steps + active.to(torch.int32)
    return torch.clamp(t, max=max_range), steps
