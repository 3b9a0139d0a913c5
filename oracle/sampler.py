"""Start-body generation oracle (TEST INFRASTRUCTURE): per-body CPU restatement of CrowdMotion.gen_init_body
(exp_GAMMAPrimitive/utils/environments.py:1041-1131) on top of the LBS oracle. pytorch3d is absent, so its
axis_angle_to_matrix / matrix_to_axis_angle are replaced by scipy's Rotation (same rotations; parity unpinned for the
third-party conversion, compared as rotation matrices in the tests). Pinned: tests/golden/gen_sampler_golden.py runs the
reference's own gen_init_body on its subseq_00343 motion seed (body model / pytorch3d served by the oracle / scipy) and
tests/test_oracle_golden.py::test_start_body_oracle_matches_reference_sampler reproduces its outputs."""
import numpy as np
import torch
from scipy.spatial.transform import Rotation


def gen_init_body(parser, start, target, betas, body_pose, global_orient, transl, yaw):
    """parser: SMPLXParserOracle; betas [10], body_pose [2,63], global_orient [2,3], transl [2,3], yaw scalar (the
    reference draws it uniformly, :1100). Returns dict(transl, global_orient_matrix, wpath)."""
    f32 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32)
    betas, body_pose, global_orient, transl = f32(betas), f32(body_pose), f32(global_orient), f32(transl)
    wpath = torch.stack([f32(start), f32(target)])

    def joints(tr, glo_aa):
        xb = torch.zeros(2, 93)
        xb[:, 0:3], xb[:, 3:6], xb[:, 6:69] = tr, glo_aa, body_pose
        return parser.forward_smplx(betas, "male", xb, "raw").joints

    j = joints(transl, global_orient)
    x_axis = j[:, 2, :] - j[:, 1, :]
    x_axis[:, -1] = 0
    x_axis = x_axis / torch.norm(x_axis, dim=-1, keepdim=True).clip(min=1e-12)
    z_axis = torch.tensor([[0.0, 0.0, 1.0]]).repeat(2, 1)
    y_axis = torch.cross(z_axis, x_axis, dim=-1)
    b_ori = y_axis[0] / torch.linalg.norm(y_axis[0])
    target_ori = wpath[1] - wpath[0]
    target_ori = target_ori / torch.linalg.norm(target_ori)
    v = torch.cross(b_ori, target_ori, dim=-1)
    c = torch.dot(b_ori, target_ori)
    s = torch.linalg.norm(v)
    kmat = torch.tensor([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    target_rot = torch.eye(3) + kmat + (kmat @ kmat) * ((1 - c) / (s ** 2))
    pelvis_zero = parser.forward_smplx(betas, "male", torch.zeros(1, 93), "raw").joints[:1, 0, :]
    aa2m = lambda aa: torch.as_tensor(Rotation.from_rotvec(aa.double().numpy()).as_matrix(), dtype=torch.float32)
    m2aa = lambda m: torch.as_tensor(Rotation.from_matrix(m.double().numpy()).as_rotvec(), dtype=torch.float32)
    rot = torch.einsum("ij,bjk->bik", target_rot, aa2m(global_orient))
    tr = torch.einsum("ij,bj->bi", target_rot, pelvis_zero + transl) - pelvis_zero
    rz = torch.as_tensor(Rotation.from_euler("z", float(yaw)).as_matrix(), dtype=torch.float32)
    rot = torch.einsum("ij,bjk->bik", rz, rot)
    tr = torch.einsum("ij,bj->bi", rz, pelvis_zero + tr) - pelvis_zero
    j = joints(tr, m2aa(rot))
    fix = torch.stack([j[0, 0, 0], j[0, 0, 1], j[0, :, 2].amin()])
    tr = tr - fix + wpath[:1]
    j = joints(tr, m2aa(rot))
    wpath[0] = j[0, 0, :]
    wpath[1, 2] = wpath[0, 2]
    return dict(transl=tr, global_orient_matrix=rot, wpath=wpath, joints=j)


def canonicalize_subsequence(parser, parser_cmu, betas, transl_all, pose_all, start_frame, end_frame, downsample_rate=3):
    """utils_canonicalize_samp.py:123-187 on the LBS oracle (torch branch of update_transl_glorot for the frame change; the
    reference script uses scipy's Rotation in float64 for the same rotation). parser / parser_cmu: SMPLXParserOracle with
    the ssm2_67 / cmu_41 marker sets. Returns the primitive dict (numpy), or None when the recording is too short."""
    if transl_all.shape[0] <= end_frame:
        return None
    f32 = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32)
    tr = f32(transl_all[start_frame:end_frame:downsample_rate])
    po = f32(pose_all[start_frame:end_frame:downsample_rate])
    be = f32(betas[:10]).reshape(1, 10)
    T = tr.shape[0]
    xb = torch.cat([tr, po[:, :66], torch.zeros(T, 24)], dim=1)
    R, Tt = parser.get_new_coordinate(be, "male", xb[:1])
    xn = parser.update_transl_glorot(R.repeat(T, 1, 1), Tt.repeat(T, 1, 1), be, "male", xb)
    poses = po.clone()
    poses[:, :3] = xn[:, 3:6]
    return {"transf_rotmat": R[0].numpy(), "transf_transl": Tt[0].numpy(), "trans": xn[:, :3].numpy(), "poses": poses.numpy(),
            "joints": parser.get_jts(be, "male", xn).numpy(),
            "marker_ssm2_67": parser.get_markers(be, "male", xn).reshape(T, -1, 3).numpy(),
            "marker_cmu_41": parser_cmu.get_markers(be, "male", xn).reshape(T, -1, 3).numpy()}
