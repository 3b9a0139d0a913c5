"""torchgeometry==0.1.2 rotation conversions, restated (TEST INFRASTRUCTURE - see oracle/__init__.py).

The reference calls these at motion/models/baseops.py:139,161,171,587,590 (aa2cont, rotmat2aa,
aa2rotmat, update_transl_glorot). torchgeometry is not vendored in /root/reference and not
installed here => parity unpinned; this file follows the published 0.1.2 ``conversions.py``
algorithm with the bool-mask arithmetic written as it behaves in the
authors' patched environment (masks as 0/1 floats).
"""
import torch


def angle_axis_to_rotation_matrix(angle_axis: torch.Tensor) -> torch.Tensor:
    """[N,3] -> [N,4,4]. Normal branch when theta^2 > 1e-6 (axis = r/(theta+1e-6)), else I+[r]x."""
    aa = angle_axis
    theta2 = (aa.unsqueeze(1) @ aa.unsqueeze(2)).squeeze(1)          # [N,1] (matmul like the original)
    eps = 1e-6
    theta = torch.sqrt(theta2)
    wxyz = aa / (theta + eps)
    wx, wy, wz = torch.chunk(wxyz, 3, dim=1)
    c = torch.cos(theta)
    s = torch.sin(theta)
    k_one = 1.0
    r00 = c + wx * wx * (k_one - c)
    r10 = wz * s + wx * wy * (k_one - c)
    r20 = -wy * s + wx * wz * (k_one - c)
    r01 = wx * wy * (k_one - c) - wz * s
    r11 = c + wy * wy * (k_one - c)
    r21 = wx * s + wy * wz * (k_one - c)
    r02 = wy * s + wx * wz * (k_one - c)
    r12 = -wx * s + wy * wz * (k_one - c)
    r22 = c + wz * wz * (k_one - c)
    normal = torch.cat([r00, r01, r02, r10, r11, r12, r20, r21, r22], dim=1).view(-1, 3, 3)
    rx, ry, rz = torch.chunk(aa, 3, dim=1)
    one = torch.ones_like(rx)
    taylor = torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).view(-1, 3, 3)
    mask = (theta2 > eps).view(-1, 1, 1)
    mask_pos = mask.type_as(theta2)
    mask_neg = (~mask).type_as(theta2)
    out = torch.eye(4, dtype=aa.dtype).view(1, 4, 4).repeat(aa.shape[0], 1, 1)
    out[..., :3, :3] = mask_pos * normal + mask_neg * taylor
    return out


def rotation_matrix_to_quaternion(rotation_matrix: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """[N,3,4] -> [N,4] (w,x,y,z). Works on the TRANSPOSE of the input, 4 cases."""
    rmat_t = torch.transpose(rotation_matrix, 1, 2)
    mask_d2 = rmat_t[:, 2, 2] < eps
    mask_d0_d1 = rmat_t[:, 0, 0] > rmat_t[:, 1, 1]
    mask_d0_nd1 = rmat_t[:, 0, 0] < -rmat_t[:, 1, 1]

    t0 = 1 + rmat_t[:, 0, 0] - rmat_t[:, 1, 1] - rmat_t[:, 2, 2]
    q0 = torch.stack([rmat_t[:, 1, 2] - rmat_t[:, 2, 1], t0,
                      rmat_t[:, 0, 1] + rmat_t[:, 1, 0], rmat_t[:, 2, 0] + rmat_t[:, 0, 2]], -1)
    t1 = 1 - rmat_t[:, 0, 0] + rmat_t[:, 1, 1] - rmat_t[:, 2, 2]
    q1 = torch.stack([rmat_t[:, 2, 0] - rmat_t[:, 0, 2], rmat_t[:, 0, 1] + rmat_t[:, 1, 0],
                      t1, rmat_t[:, 1, 2] + rmat_t[:, 2, 1]], -1)
    t2 = 1 - rmat_t[:, 0, 0] - rmat_t[:, 1, 1] + rmat_t[:, 2, 2]
    q2 = torch.stack([rmat_t[:, 0, 1] - rmat_t[:, 1, 0], rmat_t[:, 2, 0] + rmat_t[:, 0, 2],
                      rmat_t[:, 1, 2] + rmat_t[:, 2, 1], t2], -1)
    t3 = 1 + rmat_t[:, 0, 0] + rmat_t[:, 1, 1] + rmat_t[:, 2, 2]
    q3 = torch.stack([t3, rmat_t[:, 1, 2] - rmat_t[:, 2, 1],
                      rmat_t[:, 2, 0] - rmat_t[:, 0, 2], rmat_t[:, 0, 1] - rmat_t[:, 1, 0]], -1)

    f = lambda m: m.view(-1, 1).type_as(q0)
    mask_c0 = f(mask_d2 & mask_d0_d1)
    mask_c1 = f(mask_d2 & ~mask_d0_d1)
    mask_c2 = f(~mask_d2 & mask_d0_nd1)
    mask_c3 = f(~mask_d2 & ~mask_d0_nd1)
    q = q0 * mask_c0 + q1 * mask_c1 + q2 * mask_c2 + q3 * mask_c3
    rep = lambda t: t.repeat(4, 1).t()
    q = q / torch.sqrt(rep(t0) * mask_c0 + rep(t1) * mask_c1 + rep(t2) * mask_c2 + rep(t3) * mask_c3)
    q = q * 0.5
    return q


def quaternion_to_angle_axis(quaternion: torch.Tensor) -> torch.Tensor:
    q1 = quaternion[..., 1]
    q2 = quaternion[..., 2]
    q3 = quaternion[..., 3]
    sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3
    sin_theta = torch.sqrt(sin_squared_theta)
    cos_theta = quaternion[..., 0]
    two_theta = 2.0 * torch.where(cos_theta < 0.0,
                                  torch.atan2(-sin_theta, -cos_theta),
                                  torch.atan2(sin_theta, cos_theta))
    k_pos = two_theta / sin_theta
    k_neg = 2.0 * torch.ones_like(sin_theta)
    k = torch.where(sin_squared_theta > 0.0, k_pos, k_neg)
    angle_axis = torch.zeros_like(quaternion)[..., :3]
    angle_axis[..., 0] += q1 * k
    angle_axis[..., 1] += q2 * k
    angle_axis[..., 2] += q3 * k
    return angle_axis


def rotation_matrix_to_angle_axis(rotation_matrix: torch.Tensor) -> torch.Tensor:
    """[N,3,4] -> [N,3]."""
    return quaternion_to_angle_axis(rotation_matrix_to_quaternion(rotation_matrix))


def cont2rotmat(data_in: torch.Tensor) -> torch.Tensor:
    """6-D continuous rotation -> rotation matrix; follows baseops.py:120-130 (RotConverter.cont2rotmat):
    input viewed as [-1,3,2]; Gram-Schmidt on the two columns; third column = b1 x b2."""
    import torch.nn.functional as F
    x = data_in.contiguous().view(-1, 3, 2)
    b1 = F.normalize(x[:, :, 0], dim=1)
    dot = torch.sum(b1 * x[:, :, 1], dim=1, keepdim=True)
    b2 = F.normalize(x[:, :, 1] - dot * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=-1)


def rotmat2aa(data_in: torch.Tensor) -> torch.Tensor:
    """baseops.py:155-162: pad [.,3,3] to [.,3,4] then tgm.rotation_matrix_to_angle_axis."""
    import torch.nn.functional as F
    return rotation_matrix_to_angle_axis(F.pad(data_in.reshape(-1, 3, 3), [0, 1])).view(-1, 3).contiguous()


def cont2aa(data_in: torch.Tensor) -> torch.Tensor:
    """baseops.py:144-152: [N,nj,6] -> [N,nj,3]."""
    n = data_in.shape[0]
    return rotmat2aa(cont2rotmat(data_in).view(n, -1, 9)).contiguous().view(n, -1, 3)
