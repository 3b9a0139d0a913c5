"""Start-body generation and the sampler-dict contract of the reset path - mirror of
exp_GAMMAPrimitive/utils/environments.py:1041-1157 (CrowdMotion.gen_init_body / next_body) and of the dict the reference's
samplers hand to CrowdEnv.reset (crowd_env_2f.py:320-415, _canonicalize_2frame :615-644):

    {'gender', 'motion_seed': {'betas'[2,10], 'body_pose'[2,63], 'global_orient'[2,3], 'transl'[2,3]}, 'betas'[10],
     'wpath'[2,3], 'scene_path', 'navmesh', 'navmesh_path', 'floor_height'}

gen_init_body rotates a 2-frame motion seed so that the body faces its target (Rodrigues rotation between the body's
forward axis and the start->target direction, :1076-1097), optionally yaws it a little (:1099-1106) and snaps it to the
start point with the lowest joint on the floor (:1108-1115). Batched over n bodies here; the SMPL-X joints come from the
CUDA LBS operator (LbsModel), everything else is a handful of 3x3 products on the device.

pytorch3d's axis_angle_to_matrix / matrix_to_axis_angle / euler_angles_to_matrix are restated (the package is not a
dependency): Rodrigues' formula and the quaternion log map with a non-negative real part - the same rotations; the
axis-angle vector can differ from pytorch3d's only by the 2*pi branch at angle pi.
"""
from __future__ import annotations

import numpy as np
import torch


def axis_angle_to_matrix(aa: torch.Tensor) -> torch.Tensor:
    """[...,3] -> [...,3,3] (pytorch3d.transforms.axis_angle_to_matrix)."""
    ang = aa.norm(dim=-1, keepdim=True)
    small = ang < 1e-8
    ax = aa / torch.where(small, torch.ones_like(ang), ang)
    x, y, z = ax.unbind(-1)
    zero = torch.zeros_like(x)
    K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], dim=-1).reshape(aa.shape[:-1] + (3, 3))
    s, c = torch.sin(ang)[..., None], torch.cos(ang)[..., None]
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device).expand(K.shape)
    return eye + s * K + (1 - c) * (K @ K)


def matrix_to_axis_angle(R: torch.Tensor) -> torch.Tensor:
    """[...,3,3] -> [...,3] through the unit quaternion with w >= 0 (pytorch3d.transforms.matrix_to_axis_angle)."""
    m = R.reshape(-1, 3, 3)
    tr = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    # Shepperd: pick the largest of (w, x, y, z) as pivot for a well-conditioned extraction
    q_abs = torch.sqrt(torch.clamp(torch.stack([1 + tr, 1 + m[:, 0, 0] - m[:, 1, 1] - m[:, 2, 2],
                                                1 - m[:, 0, 0] + m[:, 1, 1] - m[:, 2, 2],
                                                1 - m[:, 0, 0] - m[:, 1, 1] + m[:, 2, 2]], dim=1), min=0))
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m[:, 2, 1] - m[:, 1, 2], m[:, 0, 2] - m[:, 2, 0], m[:, 1, 0] - m[:, 0, 1]], dim=1),
        torch.stack([m[:, 2, 1] - m[:, 1, 2], q_abs[:, 1] ** 2, m[:, 1, 0] + m[:, 0, 1], m[:, 0, 2] + m[:, 2, 0]], dim=1),
        torch.stack([m[:, 0, 2] - m[:, 2, 0], m[:, 1, 0] + m[:, 0, 1], q_abs[:, 2] ** 2, m[:, 1, 2] + m[:, 2, 1]], dim=1),
        torch.stack([m[:, 1, 0] - m[:, 0, 1], m[:, 2, 0] + m[:, 0, 2], m[:, 2, 1] + m[:, 1, 2], q_abs[:, 3] ** 2], dim=1),
    ], dim=1)                                                           # [n,4 pivots,4]
    cand = cand / (2.0 * q_abs[:, :, None].clamp(min=0.1))
    pick = q_abs.argmax(dim=1)
    q = cand[torch.arange(m.shape[0], device=m.device), pick]
    q = torch.where(q[:, :1] < 0, -q, q)
    vn = q[:, 1:].norm(dim=1, keepdim=True)
    ang = 2 * torch.atan2(vn, q[:, :1])
    scale = torch.where(vn < 1e-8, torch.full_like(vn, 2.0), ang / vn.clamp(min=1e-12))
    return (q[:, 1:] * scale).reshape(R.shape[:-2] + (3,))


def rotation_between(b: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """Rotation taking unit vector b onto unit vector t, [n,3] x [n,3] -> [n,3,3]: I + K + K^2 (1 - c) / s^2 with
    v = b x t, c = b.t, s = |v| (environments.py:1087-1091; singular for parallel vectors like the reference)."""
    v = torch.cross(b, t, dim=-1)
    c = (b * t).sum(-1)
    s = v.norm(dim=-1)
    zero = torch.zeros_like(c)
    K = torch.stack([zero, -v[:, 2], v[:, 1], v[:, 2], zero, -v[:, 0], -v[:, 1], v[:, 0], zero], dim=-1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=b.dtype, device=b.device).expand_as(K)
    return eye + K + (K @ K) * ((1 - c) / (s ** 2))[:, None, None]


def rot_z(theta: torch.Tensor) -> torch.Tensor:
    """euler_angles_to_matrix([0, 0, theta], 'XYZ') = Rz(theta), [n] -> [n,3,3]."""
    c, s, z, o = torch.cos(theta), torch.sin(theta), torch.zeros_like(theta), torch.ones_like(theta)
    return torch.stack([c, -s, z, s, c, z, z, z, o], dim=-1).reshape(-1, 3, 3)


class CrowdMotionSampler:
    """``CrowdMotion`` of the reference (environments.py:1007-1157) over a list of 2-frame motion seeds.

    motion_seeds: dict of arrays {'poses'[T,165 or 66], 'trans'[T,3], 'betas'[10]} (the reference's locomotion .npz
    schema, utils_canonicalize_samp.py) or a list of them. ``next_body`` keeps the reference's calling convention
    (one dict per agent); ``gen_init_bodies`` is the batched form the vector envs use."""

    def __init__(self, lbs_model, device, motion_seeds, seed: int = 0, yaw_jitter: float = 0.2):
        self.lbs, self.dev = lbs_model, torch.device(device)
        self.seeds = motion_seeds if isinstance(motion_seeds, (list, tuple)) else [motion_seeds]
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed)
        self.yaw_jitter = yaw_jitter

    def seed(self, s: int):
        self.gen.manual_seed(int(s))

    def _draw_seeds(self, n: int, fixed_seed: bool):
        betas, pose, glo, tr = [], [], [], []
        for _ in range(n):
            d = self.seeds[0] if fixed_seed else self.seeds[int(torch.randint(0, len(self.seeds), (1,), generator=self.gen))]
            f = int(torch.randint(0, len(d["poses"]) - 1, (1,), generator=self.gen))          # :1051
            betas.append(np.asarray(d["betas"], np.float32).reshape(10))
            pose.append(np.asarray(d["poses"][f:f + 2, 3:66], np.float32))
            glo.append(np.asarray(d["poses"][f:f + 2, :3], np.float32))
            tr.append(np.asarray(d["trans"][f:f + 2], np.float32))
        t = lambda x: torch.as_tensor(np.stack(x), device=self.dev)
        return t(betas), t(pose), t(glo), t(tr)

    def _joints(self, transl, glorot, pose, betas):
        """SMPL-X joints [n,2,127,3] of 2-frame bodies (hand PCA zero, like bm(**motion_seed_dict))."""
        n = transl.shape[0]
        xb = torch.zeros(n, 2, 93, device=self.dev)
        xb[:, :, 0:3], xb[:, :, 3:6], xb[:, :, 6:69] = transl, glorot, pose
        j = self.lbs.forward(xb.reshape(n * 2, 93), betas.repeat_interleave(2, 0))[1]
        return j.reshape(n, 2, -1, 3)

    def gen_init_bodies(self, start, target, fixed_seed: bool = False, motion_seed=None, yaw=None):
        """start, target [n,3] -> dict of batched tensors: transl / global_orient / body_pose [n,2,*], betas [n,10],
        wpath [n,2,3]. motion_seed = (betas [n,10], body_pose [n,2,63], global_orient [n,2,3], transl [n,2,3]) and
        yaw [n] override the random draws (tests)."""
        dev = self.dev
        start = torch.as_tensor(start, dtype=torch.float32, device=dev).reshape(-1, 3)
        target = torch.as_tensor(target, dtype=torch.float32, device=dev).reshape(-1, 3)
        n = start.shape[0]
        betas, pose, glo, tr = motion_seed if motion_seed is not None else self._draw_seeds(n, fixed_seed)
        betas, pose, glo, tr = [torch.as_tensor(x, dtype=torch.float32, device=dev) for x in (betas, pose, glo, tr)]
        wpath = torch.stack([start, target], dim=1).clone()
        # rotate the body to face the target (:1076-1097)
        j = self._joints(tr, glo, pose, betas)
        x_axis = j[:, :, 2] - j[:, :, 1]
        x_axis[..., 2] = 0
        x_axis = x_axis / x_axis.norm(dim=-1, keepdim=True).clip(min=1e-12)
        z_axis = torch.tensor([0.0, 0.0, 1.0], device=dev).expand_as(x_axis)
        b_ori = torch.cross(z_axis, x_axis, dim=-1)[:, 0]
        b_ori = b_ori / b_ori.norm(dim=-1, keepdim=True)
        t_ori = wpath[:, 1] - wpath[:, 0]
        t_ori = t_ori / t_ori.norm(dim=-1, keepdim=True)
        target_rot = rotation_between(b_ori, t_ori)[:, None]                              # [n,1,3,3]
        zero_xb = torch.zeros(n, 93, device=dev)
        pelvis_zero = self.lbs.forward(zero_xb, betas)[1][:, 0][:, None]                  # [n,1,3]
        glo = matrix_to_axis_angle(target_rot @ axis_angle_to_matrix(glo))
        tr = torch.einsum("bij,btj->bti", target_rot[:, 0], pelvis_zero + tr) - pelvis_zero
        # slightly rotate around z (:1099-1106)
        if yaw is None:
            yaw = (torch.rand(n, generator=self.gen) * 2 - 1).to(dev) * np.pi * 2 * self.yaw_jitter
        rz = rot_z(torch.as_tensor(yaw, dtype=torch.float32, device=dev))[:, None]
        glo = matrix_to_axis_angle(rz @ axis_angle_to_matrix(glo))
        tr = torch.einsum("bij,btj->bti", rz[:, 0], pelvis_zero + tr) - pelvis_zero
        # pelvis above the start point, lowest joint of frame 0 on the floor (:1108-1115)
        j = self._joints(tr, glo, pose, betas)
        fix = torch.stack([j[:, 0, 0, 0], j[:, 0, 0, 1], j[:, 0, :, 2].amin(dim=1)], dim=1)
        tr = tr - fix[:, None] + wpath[:, :1]
        j = self._joints(tr, glo, pose, betas)
        wpath[:, 0] = j[:, 0, 0]
        wpath[:, 1, 2] = wpath[:, 0, 2]
        return dict(transl=tr, global_orient=glo, body_pose=pose, betas=betas, wpath=wpath)

    def gen_init_body(self, start, target, fixed_seed: bool = False, id=None):
        """One reference-format sampler dict (environments.py:1117-1131)."""
        b = self.gen_init_bodies(np.asarray(start, np.float32)[None], np.asarray(target, np.float32)[None], fixed_seed)
        return batched_to_dicts(b)[0]

    def next_body(self, start_target=None, fixed_seed: bool = False, num_agents: int = 2):
        """environments.py:1134-1157."""
        if num_agents == 1:
            s, t = start_target
            return self.gen_init_body(s, t, fixed_seed)
        return tuple(self.gen_init_body(s, t, fixed_seed) for s, t in start_target[:num_agents])


def batched_to_dicts(b: dict):
    """Batched gen_init_bodies output -> list of reference-format sampler dicts."""
    out = []
    for i in range(b["transl"].shape[0]):
        ms = {"betas": b["betas"][i][None].repeat(2, 1), "body_pose": b["body_pose"][i], "global_orient": b["global_orient"][i],
              "transl": b["transl"][i]}
        out.append({"gender": "male", "motion_seed": ms, "betas": b["betas"][i], "wpath": b["wpath"][i],
                    "scene_path": "data/floor.ply", "navmesh": None, "navmesh_path": None, "floor_height": 0})
    return out


def sampler_dicts_to_candidates(dicts, device):
    """Reference-format sampler dicts (one per env) -> the reset inputs of the vector env: world_params [n,2,93] =
    [transl, global_orient, body_pose, 24 zeros] (_canonicalize_2frame, crowd_env_2f.py:617-626), goals [n,3] =
    wpath[-1], betas [n,10]."""
    dev = torch.device(device)
    if isinstance(dicts, dict):
        dicts = [dicts]
    n = len(dicts)
    wp = torch.zeros(n, 2, 93, device=dev)
    goals, betas = torch.zeros(n, 3, device=dev), torch.zeros(n, 10, device=dev)
    for i, d in enumerate(dicts):
        ms = d["motion_seed"]
        t = lambda x: torch.as_tensor(x, dtype=torch.float32, device=dev)
        wp[i, :, 0:3], wp[i, :, 3:6], wp[i, :, 6:69] = t(ms["transl"]), t(ms["global_orient"]), t(ms["body_pose"])
        goals[i] = t(d["wpath"])[-1]
        betas[i] = t(d["betas"]).reshape(-1)[:10]
    return wp, goals, betas
