"""smplx==0.1.28 SMPLX.forward / lbs, restated in PyTorch-CPU (TEST INFRASTRUCTURE).

Reference call sites: motion/models/baseops.py:291-320 (model creation: num_pca_comps=12,
flat_hand_mean=False default, expression/jaw/eye poses default zeros), :382 (forward),
:529 (calc_calibrate_offset). The in-tree cross-check for the lbs call pattern is
experiments/HOOD/utils/lbs.py:7-46,86-124. smplx itself is not vendored => parity unpinned;
the steps below follow the published algorithm.
"""
from types import SimpleNamespace

import numpy as np
import torch


def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    """smplx.lbs.batch_rodrigues: angle = ||r + 1e-8||, axis = r/angle, R = I + sin K + (1-cos) K^2."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.unsqueeze(torch.cos(angle), dim=1)
    sin = torch.unsqueeze(torch.sin(angle), dim=1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view((n, 3, 3))
    ident = torch.eye(3, dtype=rot_vecs.dtype).unsqueeze(dim=0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    """smplx.lbs.batch_rigid_transform: chain G_j = G_parent(j) [R_j | J_j - J_parent]; returns
    posed joints and the relative transforms A_j (translation minus G_j J_j)."""
    import torch.nn.functional as F
    B, J = joints.shape[:2]
    joints = torch.unsqueeze(joints, dim=-1)
    rel_joints = joints.clone()
    rel_joints[:, 1:] -= joints[:, parents[1:]]
    tm = torch.cat([F.pad(rot_mats.reshape(-1, 3, 3), [0, 0, 0, 1]),
                    F.pad(rel_joints.reshape(-1, 3, 1), [0, 0, 0, 1], value=1)], dim=2).reshape(-1, J, 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_homogen = F.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - F.pad(torch.matmul(transforms, joints_homogen), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


class SMPLXOracle:
    """Holds SMPLX_KEYS arrays (egogen_b200.assets) as torch-CPU float32 tensors and evaluates
    the smplx forward pass exactly in the order smplx does."""

    def __init__(self, model: dict):
        t = lambda k, dt=torch.float32: torch.as_tensor(np.asarray(model[k]), dtype=dt)
        self.v_template = t("v_template")
        self.shapedirs = t("shapedirs")            # [V,3,20]
        self.posedirs = t("posedirs")              # [486, V*3]
        self.J_regressor = t("J_regressor")        # [55,V]
        self.parents = t("parents", torch.long)
        self.lbs_weights = t("lbs_weights")        # [V,55]
        self.hand_comp_l = t("hand_comp_l")
        self.hand_comp_r = t("hand_comp_r")
        self.pose_mean = t("pose_mean")
        self.extra_vids = t("extra_vids", torch.long)
        self.faces = t("faces", torch.long)
        self.lmk_faces_idx = t("lmk_faces_idx", torch.long)
        self.lmk_bary = t("lmk_bary")

    def forward(self, betas, global_orient, body_pose, left_hand_pose, right_hand_pose, transl,
                expression=None):
        """All inputs [N,*] float32. Returns SimpleNamespace(vertices [N,V,3], joints [N,127,3])."""
        N = global_orient.shape[0]
        z3 = torch.zeros(N, 3)
        if expression is None:
            expression = torch.zeros(N, 10)
        lh = torch.einsum("bi,ij->bj", left_hand_pose, self.hand_comp_l)
        rh = torch.einsum("bi,ij->bj", right_hand_pose, self.hand_comp_r)
        full_pose = torch.cat([global_orient, body_pose, z3, z3, z3, lh, rh], dim=1)
        full_pose = full_pose + self.pose_mean
        shape_components = torch.cat([betas, expression], dim=-1)
        # lbs()
        v_shaped = self.v_template + torch.einsum("bl,mkl->bmk", shape_components, self.shapedirs)
        J = torch.einsum("bik,ji->bjk", v_shaped, self.J_regressor)
        ident = torch.eye(3)
        rot_mats = batch_rodrigues(full_pose.view(-1, 3)).view(N, -1, 3, 3)
        pose_feature = (rot_mats[:, 1:, :, :] - ident).view(N, -1)
        pose_offsets = torch.matmul(pose_feature, self.posedirs).view(N, -1, 3)
        v_posed = pose_offsets + v_shaped
        J_transformed, A = batch_rigid_transform(rot_mats, J, self.parents)
        W = self.lbs_weights.unsqueeze(0).expand(N, -1, -1)
        nj = self.J_regressor.shape[0]
        T = torch.matmul(W, A.view(N, nj, 16)).view(N, -1, 4, 4)
        v_homo = torch.matmul(T, torch.cat([v_posed, torch.ones(N, v_posed.shape[1], 1)], dim=2).unsqueeze(-1))
        verts = v_homo[:, :, :3, 0]
        # landmarks (static 51), vertex joints (21), concat
        lmk_faces = self.faces[self.lmk_faces_idx]                  # [51,3]
        lmk_vertices = verts[:, lmk_faces]                          # [N,51,3,3]
        landmarks = torch.einsum("blfi,lf->bli", lmk_vertices, self.lmk_bary)
        joints = torch.cat([J_transformed, verts[:, self.extra_vids], landmarks], dim=1)
        joints = joints + transl.unsqueeze(1)
        verts = verts + transl.unsqueeze(1)
        return SimpleNamespace(vertices=verts, joints=joints, full_pose=full_pose, v_shaped=v_shaped, A=A)


class SMPLXParserOracle:
    """Restatement of SMPLXParser's torch branch (baseops.py:271-598), male/female models given
    as SMPLX_KEYS dicts. Tensors in/out on CPU."""

    def __init__(self, model_male: dict, model_female: dict = None, marker: list = None):
        from . import tgm  # noqa
        self.bm_male = SMPLXOracle(model_male)
        self.bm_female = SMPLXOracle(model_female) if model_female is not None else self.bm_male
        self.marker = marker

    def _bm(self, gender):
        return self.bm_male if gender == "male" else self.bm_female

    def forward_smplx(self, betas, gender, xb, output_type="markers"):
        """baseops.py:338-398 (to_numpy=False)."""
        n = xb.shape[0]
        out = self._bm(gender).forward(
            betas=betas.reshape(-1, 10).repeat(n, 1) if betas.numel() == 10 else betas.reshape(n, 10),
            global_orient=xb[:, 3:6], body_pose=xb[:, 6:69],
            left_hand_pose=xb[:, 69:81], right_hand_pose=xb[:, 81:], transl=xb[:, :3])
        if output_type == "markers":
            return out.vertices[:, self.marker, :]
        if output_type == "joints":
            return out.joints[:, :22]
        if output_type == "all_joints":
            return out.joints
        if output_type == "vertices":
            return out.vertices
        if output_type == "raw":
            return out
        raise NotImplementedError("other output types are not supported")

    def get_jts(self, betas, gender, xb):
        return self.forward_smplx(betas, gender, xb, "joints")

    def get_all_jts(self, betas, gender, xb):
        return self.forward_smplx(betas, gender, xb, "all_joints")

    def get_markers(self, betas, gender, xb):
        return self.forward_smplx(betas, gender, xb, "markers")

    @staticmethod
    def new_coordinate_from_joints(jts):
        """CanonicalCoordinateExtractor.get_new_coordinate_torch, baseops.py:214-225."""
        x_axis = jts[:, 2, :] - jts[:, 1, :]
        x_axis = x_axis.clone()
        x_axis[:, -1] = 0
        x_axis = x_axis / torch.norm(x_axis, dim=-1, keepdim=True)
        z_axis = torch.tensor([[0.0, 0.0, 1.0]]).repeat(x_axis.shape[0], 1)
        y_axis = torch.cross(z_axis, x_axis, dim=-1)
        y_axis = y_axis / torch.norm(y_axis, dim=-1, keepdim=True)
        return torch.stack([x_axis, y_axis, z_axis], dim=-1), jts[:, :1]

    def get_new_coordinate(self, betas, gender, xb):
        """baseops.py:465-490."""
        return self.new_coordinate_from_joints(self.get_jts(betas, gender, xb))

    def calc_calibrate_offset(self, gender, betas, body_pose):
        """baseops.py:494-534: pelvis of the body with zero transl / global_orient / hands."""
        n = body_pose.shape[0]
        z = torch.zeros
        out = self._bm(gender).forward(betas=betas.reshape(-1, 10).repeat(n, 1), global_orient=z(n, 3),
                                       body_pose=body_pose, left_hand_pose=z(n, 12),
                                       right_hand_pose=z(n, 12), transl=z(n, 3))
        return out.joints[:, 0, :]

    def update_transl_glorot(self, transf_rotmat, transf_transl, betas, gender, xb):
        """baseops.py:537-598, torch branch, inplace=False."""
        import torch.nn.functional as F
        from . import tgm
        delta_T = self.calc_calibrate_offset(gender, betas, xb[:, 6:69])
        transl = xb[:, :3]
        glorot = xb[:, 3:6]
        global_ori = tgm.angle_axis_to_rotation_matrix(glorot)[:, :3, :3]
        global_ori_new = torch.einsum("bij,bjk->bik", transf_rotmat.permute(0, 2, 1), global_ori)
        glorot = tgm.rotation_matrix_to_angle_axis(F.pad(global_ori_new, [0, 1])).view(-1, 3).contiguous()
        transl = torch.einsum("bij,bj->bi", transf_rotmat.permute(0, 2, 1),
                              transl + delta_T - transf_transl[:, 0]) - delta_T
        return torch.cat([transl, glorot, xb[:, 6:]], dim=1)
