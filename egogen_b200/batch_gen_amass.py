"""Canonicalised motion-primitive data: the `.npz` schema and its batch generator - host-side mirror of
motion/exp_GAMMAPrimitive/utils/utils_canonicalize_samp.py:123-187 (canonicalize_subsequence) and
motion/exp_GAMMAPrimitive/utils/batch_gen_amass.py:61-429 (BatchGeneratorAMASSCanonicalized).

One primitive file holds
    trans [T,3], poses [T,>=66], betas [>=10], gender, mocap_framerate (=120),
    joints [T,22,3], marker_cmu_41 [T,41,3], marker_ssm2_67 [T,67,3], transf_rotmat [3,3], transf_transl [1,3]
all expressed in the canonical frame of the first frame. Everything that touches the body model (the canonical frame, the
pelvis offset, joints / markers, the noise-augmented batches) runs through the CUDA library (SMPLXParser mirror); file
parsing, filtering, shuffling and batching are host logic.
"""
from __future__ import annotations

import glob
import os
import random

import numpy as np
import torch

PRIMITIVE_KEYS = ("trans", "poses", "betas", "gender", "mocap_framerate", "joints", "marker_cmu_41", "marker_ssm2_67",
                  "transf_rotmat", "transf_transl")


def _xb(transl, pose66):
    """[T,93] body vector (hands at the mean pose, as in the canonicalisation scripts)."""
    t = transl.shape[0]
    return np.concatenate([transl, pose66, np.zeros((t, 24), dtype=transl.dtype)], axis=1).astype(np.float32)


def canonicalize_subsequence(parsers, betas, transl_all, pose_all, start_frame, end_frame, gender="male", fps=120,
                             downsample_rate=3):
    """utils_canonicalize_samp.py:123-187. `parsers` = {'ssm2_67': SMPLXParser, 'cmu_41': SMPLXParser} (one marker
    placement per parser). Returns the primitive dict, or None when the recording is too short."""
    assert fps == 120
    if transl_all.shape[0] <= end_frame:
        return None
    p67 = parsers["ssm2_67"]
    betas = np.asarray(betas, dtype=np.float32)
    transl = np.array(transl_all[start_frame:end_frame:downsample_rate], dtype=np.float32)
    pose = np.array(pose_all[start_frame:end_frame:downsample_rate], dtype=np.float32)
    # frame of the first body, then transl / global_orient re-expressed in it (offset-compensated)
    R, T = p67.get_new_coordinate(betas[:10], gender, _xb(transl[:1], pose[:1, :66]), to_numpy=True)
    t = transl.shape[0]
    xb = p67.update_transl_glorot(np.repeat(R, t, 0), np.repeat(T, t, 0), betas[:10], gender, _xb(transl, pose[:, :66]),
                                  to_numpy=True, inplace=False)
    pose[:, :3] = xb[:, 3:6]
    out = {"transf_rotmat": R[0], "transf_transl": T[0], "trans": xb[:, :3].copy(), "poses": pose, "betas": betas,
           "gender": gender, "mocap_framerate": int(fps)}
    out["joints"] = p67.get_jts(betas[:10], gender, xb)
    out["marker_ssm2_67"] = p67.get_markers(betas[:10], gender, xb)
    out["marker_cmu_41"] = parsers["cmu_41"].get_markers(betas[:10], gender, xb)
    return out


def save_primitive(path, data_out):
    """np.savez with the reference's keys (utils_canonicalize_samp.py:262-287)."""
    missing = [k for k in PRIMITIVE_KEYS if k not in data_out]
    if missing:
        raise KeyError(f"primitive is missing {missing}")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, **{k: data_out[k] for k in PRIMITIVE_KEYS})


def body_feature(body_repr, transl, pose, joints, body_cmu_41, body_ssm2_67, marker2tarloc_n):
    """The per-frame feature selected by `body_repr` (batch_gen_amass.py:193-214). NB the reference reshapes the 41
    CMU markers with 67*3 inside get_rec_list (:199, a latent bug that raises for cmu_41); here cmu_41 gives [T,123]
    as in next_sequence (:306)."""
    if body_repr == "smpl_params":
        return np.concatenate([transl, pose], axis=-1)
    if body_repr == "joints":
        return joints.reshape([-1, 22 * 3])
    if body_repr == "cmu_41":
        return body_cmu_41.reshape([-1, 41 * 3])
    if body_repr == "ssm2_67":
        return body_ssm2_67.reshape([-1, 67 * 3])
    if body_repr == "ssm2_67_marker2tarloc":
        return np.concatenate([body_ssm2_67.reshape([-1, 67 * 3]), marker2tarloc_n.reshape([-1, 67 * 3])], axis=-1)
    if body_repr == "bone_transform":
        return np.concatenate([joints, pose.reshape([-1, 22, 3])], axis=-1)
    raise NameError("[ERROR] not valid body representation. Terminate")


def get_target_feature(joints, body_ssm2_67, rotmat=np.eye(3), transl=np.zeros((1, 3))):
    """_get_target_feature (batch_gen_amass.py:270-283). Like the reference it lowers the last pelvis IN PLACE by the
    frame translation's height (joints[-1, 0, 2] -= transl[0, 2])."""
    wpath = joints[-1:] - joints
    wpath = wpath[:, 0, :2]
    wpath_n = wpath / (1e-8 + np.linalg.norm(wpath, axis=-1, keepdims=True))
    vec_to_target = body_ssm2_67[-1:] - body_ssm2_67
    target_loc = joints[-1:, 0:1]
    target_loc[:, :, -1] = target_loc[:, :, -1] - transl[None, ...][:, :, -1]
    vec_to_target_loc = target_loc - body_ssm2_67
    vec_to_target_locn = vec_to_target_loc / np.linalg.norm(vec_to_target_loc, axis=-1, keepdims=True)
    return vec_to_target, wpath_n, vec_to_target_locn


def apply_rot_noise(rot, noise):
    """batch_gen_amass.py:32-37: R(noise) R(rot) per joint, axis-angle in and out (pytorch3d conventions)."""
    from .scene_sampler import axis_angle_to_matrix, matrix_to_axis_angle
    t, d = rot.shape
    res = torch.matmul(axis_angle_to_matrix(noise.reshape(-1, 3)), axis_angle_to_matrix(rot.reshape(-1, 3)))
    return matrix_to_axis_angle(res).reshape(t, d)


class BatchGeneratorAMASSCanonicalized:
    def __init__(self, amass_data_path, amass_subset_name=None, sample_rate=3, body_repr="cmu_41", read_to_ram=True,
                 device="cuda:0", parser=None):
        self.rec_list = []
        self.index_rec = 0
        self.amass_data_path = amass_data_path
        self.amass_subset_name = amass_subset_name
        self.sample_rate = sample_rate
        self.data_list = []
        self.jts_list = []
        self.body_repr = body_repr
        self.read_to_ram = read_to_ram
        self.max_len = 200 if "x10" in amass_data_path else 20
        self.device = torch.device(device)
        self._parser = parser            # SMPLXParser (ssm2_67), built on first use by the noise-augmented batches

    # ---- iteration state ---------------------------------------------------------------------
    def _permute(self, with_jts):
        random.shuffle(self.data_list)
        idx = torch.randperm(self.data_all.shape[0])
        names = ["data_all", "pose_all", "beta_all", "transl_all", "gender_all"] + (["jts_all"] if with_jts else [])
        for n in names:
            v = getattr(self, n)
            setattr(self, n, v[idx.to(v.device)] if torch.is_tensor(v) else v[idx.numpy()])

    def reset(self):
        self.index_rec = 0
        if self.read_to_ram:
            self._permute(False)
        else:
            random.shuffle(self.rec_list)

    def reset_with_jts(self):
        self.index_rec = 0
        if self.read_to_ram:
            self._permute(True)
        else:
            random.shuffle(self.rec_list)

    def has_next_rec(self):
        return self.index_rec < (len(self.data_list) if self.read_to_ram else len(self.rec_list))

    # ---- loading -------------------------------------------------------------------------------
    def get_rec_list(self, shuffle_seed=None, to_gpu=False):
        if self.amass_subset_name is not None:
            self.rec_list = []
            for subset in self.amass_subset_name:
                self.rec_list += sorted(glob.glob(os.path.join(self.amass_data_path, subset, "*.npz")))
        else:
            self.rec_list = sorted(glob.glob(os.path.join(self.amass_data_path, "*/*.npz")))
        if shuffle_seed is not None:
            random.Random(shuffle_seed).shuffle(self.rec_list)
        else:
            random.shuffle(self.rec_list)
        if not self.read_to_ram:
            return
        self.data_list, self.jts_list = [], []
        self.pose_list, self.transl_list, self.beta_list, self.gender_list = [], [], [], []
        for rec in self.rec_list:
            with np.load(rec) as d:
                if d["mocap_framerate"] != 120:
                    continue
                sr = self.sample_rate
                pose = d["poses"][::sr, :66]
                transl = d["trans"][::sr]
                beta = d["betas"]
                gender = d["gender"].astype(str)
                if np.isnan(pose).any() or np.isinf(pose).any() or np.isnan(transl).any() or np.isinf(transl).any():
                    continue
                body_cmu_41 = d["marker_cmu_41"][::sr]
                body_ssm2_67 = d["marker_ssm2_67"][::sr]
                joints = d["joints"][::sr].reshape([-1, 22, 3])
                transf_rotmat = d["transf_rotmat"]
                transf_transl = d["transf_transl"]
            m = self.max_len
            transl, pose, body_cmu_41, body_ssm2_67, joints = transl[:m], pose[:m], body_cmu_41[:m], body_ssm2_67[:m], joints[:m]
            _, _, marker2tarloc_n = get_target_feature(joints, body_ssm2_67, transf_rotmat, transf_transl)
            self.data_list.append(body_feature(self.body_repr, transl, pose, joints, body_cmu_41, body_ssm2_67, marker2tarloc_n))
            self.jts_list.append(joints)
            self.pose_list.append(pose)
            self.beta_list.append(beta)
            self.transl_list.append(transl)
            self.gender_list.append(gender)
        if not self.data_list:
            raise FileNotFoundError(f"no usable 120 fps primitives under {self.amass_data_path}")
        self.data_all = np.stack(self.data_list, axis=0).astype(np.float32)        # [b,t,d]
        self.jts_all = np.stack(self.jts_list, axis=0).astype(np.float32)          # [b,t,22,3]
        # the reference stacks data_list here (:223), which feeds marker coordinates to the noise path as if they were
        # joint rotations; the rotations are what apply_rot_noise / the body model need
        self.pose_all = np.stack(self.pose_list, axis=0).astype(np.float32)        # [b,t,66]
        self.beta_all = np.stack(self.beta_list, axis=0).astype(np.float32)
        self.transl_all = np.stack(self.transl_list, axis=0).astype(np.float32)
        self.gender_all = np.array(self.gender_list)
        if to_gpu:
            for n in ("data_all", "jts_all", "pose_all", "transl_all", "beta_all"):
                setattr(self, n, torch.as_tensor(getattr(self, n)).to(self.device))

    # ---- batches ---------------------------------------------------------------------------------
    def _dev(self, x):
        t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x), dtype=torch.float32)
        return t.to(self.device, dtype=torch.float32)

    def _get_parser(self):
        if self._parser is None:
            from .smplx_parser import SMPLXParser
            self._parser = SMPLXParser({"n_batch": self.max_len, "device": self.device, "marker_placement": "ssm2_67"})
        return self._parser

    def next_batch(self, batch_size=64, noise=None):
        """[t,b,d]. With `noise` (std of a per-sequence axis-angle perturbation applied to every joint, constant over
        the sequence) the markers are re-generated by the body model from the perturbed poses (:240-263)."""
        if noise is None:
            batch = self.data_all[self.index_rec:self.index_rec + batch_size]
            self.index_rec += batch_size
            return self._dev(batch).permute(1, 0, 2)
        parser = self._get_parser()
        out, bb = [], 0
        while self.has_next_rec():
            if bb == batch_size:
                break
            i = self.index_rec
            gender = str(self.gender_all[i])
            transl, pose, betas = self._dev(self.transl_all[i]), self._dev(self.pose_all[i]), self._dev(self.beta_all[i])
            rot_noise = torch.normal(mean=0.0, std=float(noise), size=pose[:1].shape, device=pose.device).expand(pose.shape)
            pose = apply_rot_noise(pose, rot_noise)
            t = pose.shape[0]
            xb = torch.cat([transl, pose, torch.zeros(t, 24, device=pose.device)], dim=1)
            mk = parser.forward_smplx(betas[:10].reshape(1, 10), gender, xb, to_numpy=False, output_type="markers")
            out.append(mk.reshape(t, 67 * 3))
            self.index_rec += 1
            bb += 1
            if self.index_rec == len(self.data_list):
                break
        return torch.stack(out).permute(1, 0, 2)

    def next_batch_with_jts(self, batch_size=64, noise=None):
        d = self._dev(self.data_all[self.index_rec:self.index_rec + batch_size]).permute(1, 0, 2)
        j = self._dev(self.jts_all[self.index_rec:self.index_rec + batch_size]).permute(1, 0, 2, 3)
        self.index_rec += batch_size
        return d, j

    def next_batch_genderselection(self, batch_size=64, gender="male", batch_first=True, noise=None):
        """Same-gender batch read from the files (:348-429): [betas, body_feature, transl, glorot, thetas, joints],
        each [b,t,d] (or [t,b,d]); None when fewer than batch_size sequences are left."""
        keys = ("betas", "transl", "glorot", "thetas", "feature", "jts")
        acc = {k: [] for k in keys}
        bb = 0
        while self.has_next_rec():
            rec = self.rec_list[self.index_rec]
            if bb == batch_size:
                break
            with np.load(rec) as d:
                if str(d["gender"]) != gender:
                    self.index_rec += 1
                    continue
                sr = self.sample_rate
                transl = d["trans"][::sr]
                pose = d["poses"][::sr, :66]
                betas = np.tile(d["betas"][:10], (transl.shape[0], 1))
                body_cmu_41 = d["marker_cmu_41"][::sr]
                body_ssm2_67 = d["marker_ssm2_67"][::sr]
                joints = d["joints"][::sr].reshape([-1, 22, 3])
            _, _, marker2tarloc_n = get_target_feature(joints, body_ssm2_67)
            acc["feature"].append(body_feature(self.body_repr, transl, pose, joints, body_cmu_41, body_ssm2_67, marker2tarloc_n))
            acc["betas"].append(betas); acc["transl"].append(transl); acc["glorot"].append(pose[:, :3])
            acc["thetas"].append(pose[:, 3:]); acc["jts"].append(joints.reshape([-1, 22 * 3]))
            self.index_rec += 1
            bb += 1
            if self.index_rec == len(self.data_list):
                break
        if len(acc["betas"]) < batch_size:
            return None
        ax = 0 if batch_first else 1
        st = {k: self._dev(np.stack(v, axis=ax).astype(np.float32)) for k, v in acc.items()}
        return [st["betas"], st["feature"], st["transl"], st["glorot"], st["thetas"], st["jts"]]

    def next_sequence(self):
        """One recording with its meta information (:287-345); None for recordings with NaN / inf parameters."""
        rec = self.rec_list[self.index_rec]
        with np.load(rec) as d:
            sr = self.sample_rate
            pose = d["poses"][::sr, :66]
            transl = d["trans"][::sr]
            gender = d["gender"]
            if np.isnan(pose).any() or np.isinf(pose).any() or np.isnan(transl).any() or np.isinf(transl).any():
                return None
            betas = d["betas"][:10]
            body_cmu_41 = d["marker_cmu_41"][::sr]
            body_ssm2_67 = d["marker_ssm2_67"][::sr]
            joints = d["joints"][::sr].reshape([-1, 22, 3])
            transf_rotmat = d["transf_rotmat"]
            transf_transl = d["transf_transl"]
        _, _, marker2tarloc_n = get_target_feature(joints, body_ssm2_67)
        feat = body_feature(self.body_repr, transl, pose, joints, body_cmu_41, body_ssm2_67, marker2tarloc_n)
        self.index_rec += 1
        return {"betas": betas, "gender": gender, "transl": transl, "glorot": pose[:, :3], "poses": pose[:, 3:],
                "body_feature": feat, "transf_rotmat": transf_rotmat, "transf_transl": transf_transl,
                "pelvis_loc": joints[:, 0, :]}

    def get_all_data(self):
        return torch.as_tensor(self.data_all, dtype=torch.float32).permute(1, 0, 2)
