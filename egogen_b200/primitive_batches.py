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


def _finite(*arrays):
    return all(np.isfinite(a).all() for a in arrays)


def read_primitive(path, stride=1, max_len=None, need_120fps=False, check_finite=True):
    """One primitive file as a dict of strided (and optionally truncated) arrays, or None when it fails the filters the
    reference applies while loading (:161-171: frame rate, NaN / inf in the body parameters)."""
    with np.load(path) as f:
        if need_120fps and f["mocap_framerate"] != 120:
            return None
        rec = {"pose": f["poses"][::stride, :66], "transl": f["trans"][::stride], "betas": f["betas"],
               "gender": str(f["gender"].astype(str)), "cmu": f["marker_cmu_41"][::stride], "ssm": f["marker_ssm2_67"][::stride],
               "joints": f["joints"][::stride].reshape([-1, 22, 3]), "R": f["transf_rotmat"], "T": f["transf_transl"]}
    if check_finite and not _finite(rec["pose"], rec["transl"]):
        return None
    if max_len is not None:
        for k in ("pose", "transl", "cmu", "ssm", "joints"):
            rec[k] = rec[k][:max_len]
    return rec


class BatchGeneratorAMASSCanonicalized:
    """Same constructor arguments, attributes and batch methods as the reference class; `device` / `parser` are the
    only additions (where batches land, and the body model used by the noise-augmented batches)."""

    _IN_MEMORY = ("data_all", "pose_all", "beta_all", "transl_all", "gender_all")

    def __init__(self, amass_data_path, amass_subset_name=None, sample_rate=3, body_repr="cmu_41", read_to_ram=True,
                 device="cuda:0", parser=None):
        self.amass_data_path, self.amass_subset_name = amass_data_path, amass_subset_name
        self.sample_rate, self.body_repr, self.read_to_ram = sample_rate, body_repr, read_to_ram
        self.rec_list, self.data_list, self.jts_list = [], [], []
        self.index_rec = 0
        self.max_len = 200 if "x10" in amass_data_path else 20
        self.device = torch.device(device)
        self._parser = parser            # SMPLXParser (ssm2_67), built on first use by the noise-augmented batches

    # ---- iteration state ---------------------------------------------------------------------
    def _rewind(self, fields):
        """index back to 0 and a fresh order: one permutation applied to the named in-memory arrays (the reference's
        reset() leaves jts_all out, reset_with_jts() includes it), or a shuffle of the file list."""
        self.index_rec = 0
        if not self.read_to_ram:
            random.shuffle(self.rec_list)
            return
        random.shuffle(self.data_list)
        order = torch.randperm(self.data_all.shape[0])
        for name in fields:
            v = getattr(self, name)
            setattr(self, name, v[order.to(v.device)] if torch.is_tensor(v) else v[order.numpy()])

    def reset(self):
        self._rewind(self._IN_MEMORY)

    def reset_with_jts(self):
        self._rewind(self._IN_MEMORY + ("jts_all",))

    def has_next_rec(self):
        return self.index_rec < len(self.data_list if self.read_to_ram else self.rec_list)

    # ---- loading -------------------------------------------------------------------------------
    def _feature(self, rec, frame_R=np.eye(3), frame_T=np.zeros((1, 3))):
        _, _, to_target = get_target_feature(rec["joints"], rec["ssm"], frame_R, frame_T)
        return body_feature(self.body_repr, rec["transl"], rec["pose"], rec["joints"], rec["cmu"], rec["ssm"], to_target)

    def get_rec_list(self, shuffle_seed=None, to_gpu=False):
        root = self.amass_data_path
        patterns = [os.path.join(root, "*/*.npz")] if self.amass_subset_name is None else \
                   [os.path.join(root, sub, "*.npz") for sub in self.amass_subset_name]
        self.rec_list = [p for pat in patterns for p in sorted(glob.glob(pat))]
        (random.Random(shuffle_seed) if shuffle_seed is not None else random).shuffle(self.rec_list)
        if not self.read_to_ram:
            return
        kept = [r for r in (read_primitive(p, self.sample_rate, self.max_len, need_120fps=True) for p in self.rec_list)
                if r is not None]
        if not kept:
            raise FileNotFoundError(f"no usable 120 fps primitives under {root}")
        self.data_list = [self._feature(r, r["R"], r["T"]) for r in kept]          # also applies the pelvis quirk to joints
        self.jts_list = [r["joints"] for r in kept]
        stack = lambda key: np.stack([r[key] for r in kept], axis=0).astype(np.float32)
        self.data_all = np.stack(self.data_list, axis=0).astype(np.float32)        # [b,t,d]
        self.jts_all = stack("joints")                                              # [b,t,22,3]
        # the reference stacks data_list here (:223), which feeds marker coordinates to the noise path as if they were
        # joint rotations; the rotations are what apply_rot_noise / the body model need
        self.pose_all, self.beta_all, self.transl_all = stack("pose"), stack("betas"), stack("transl")
        self.gender_all = np.array([r["gender"] for r in kept])
        if to_gpu:
            for name in ("data_all", "jts_all", "pose_all", "transl_all", "beta_all"):
                setattr(self, name, torch.as_tensor(getattr(self, name)).to(self.device))

    # ---- batches ---------------------------------------------------------------------------------
    def _dev(self, x):
        t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x), dtype=torch.float32)
        return t.to(self.device, dtype=torch.float32)

    def _get_parser(self):
        if self._parser is None:
            from .smplx_parser import SMPLXParser
            self._parser = SMPLXParser({"n_batch": self.max_len, "device": self.device, "marker_placement": "ssm2_67"})
        return self._parser

    def _take(self, array, batch_size):
        return self._dev(array[self.index_rec:self.index_rec + batch_size])

    def _noisy_markers(self, i, noise):
        """markers [t,201] of recording i after one random rotation per joint (std `noise`, constant over the frames)
        has been composed onto its pose (:242-259)."""
        transl, pose, betas = (self._dev(a[i]) for a in (self.transl_all, self.pose_all, self.beta_all))
        jitter = torch.normal(mean=0.0, std=float(noise), size=pose[:1].shape, device=pose.device).expand(pose.shape)
        pose = apply_rot_noise(pose, jitter)
        xb = torch.cat([transl, pose, torch.zeros(pose.shape[0], 24, device=pose.device)], dim=1)
        mk = self._get_parser().forward_smplx(betas[:10].reshape(1, 10), str(self.gender_all[i]), xb, to_numpy=False,
                                              output_type="markers")
        return mk.reshape(pose.shape[0], 67 * 3)

    def next_batch(self, batch_size=64, noise=None):
        """[t,b,d]. With `noise` the markers are re-generated by the body model from perturbed poses."""
        if noise is None:
            batch = self._take(self.data_all, batch_size)
            self.index_rec += batch_size
            return batch.permute(1, 0, 2)
        last = min(self.index_rec + batch_size, len(self.data_list))
        seqs = [self._noisy_markers(i, noise) for i in range(self.index_rec, last)]
        self.index_rec = last
        return torch.stack(seqs).permute(1, 0, 2)

    def next_batch_with_jts(self, batch_size=64, noise=None):
        d, j = self._take(self.data_all, batch_size), self._take(self.jts_all, batch_size)
        self.index_rec += batch_size
        return d.permute(1, 0, 2), j.permute(1, 0, 2, 3)

    def next_batch_genderselection(self, batch_size=64, gender="male", batch_first=True, noise=None):
        """Same-gender batch read from the files (:348-429): [betas, body_feature, transl, glorot, thetas, joints],
        each [b,t,d] (or [t,b,d]); None when fewer than batch_size sequences are left."""
        picked = []
        while self.has_next_rec() and len(picked) < batch_size:
            rec = read_primitive(self.rec_list[self.index_rec], self.sample_rate, check_finite=False)   # unfiltered, as :372-381
            self.index_rec += 1
            if rec["gender"] == gender:
                picked.append(rec)
        if len(picked) < batch_size:
            return None
        cols = {"betas": lambda r: np.tile(r["betas"][:10], (r["transl"].shape[0], 1)), "feature": self._feature,
                "transl": lambda r: r["transl"], "glorot": lambda r: r["pose"][:, :3], "thetas": lambda r: r["pose"][:, 3:],
                "jts": lambda r: r["joints"].reshape([-1, 22 * 3])}
        axis = 0 if batch_first else 1
        out = {k: self._dev(np.stack([f(r) for r in picked], axis=axis).astype(np.float32)) for k, f in cols.items()}
        return [out[k] for k in ("betas", "feature", "transl", "glorot", "thetas", "jts")]

    def next_sequence(self):
        """One recording with its meta information (:287-345); None for recordings with NaN / inf parameters."""
        path = self.rec_list[self.index_rec]
        rec = read_primitive(path, self.sample_rate)
        if rec is None:
            return None
        self.index_rec += 1
        with np.load(path) as f:
            gender = f["gender"]
        return {"betas": rec["betas"][:10], "gender": gender, "transl": rec["transl"], "glorot": rec["pose"][:, :3],
                "poses": rec["pose"][:, 3:], "body_feature": self._feature(rec), "transf_rotmat": rec["R"],
                "transf_transl": rec["T"], "pelvis_loc": rec["joints"][:, 0, :]}

    def get_all_data(self):
        return torch.as_tensor(self.data_all, dtype=torch.float32).permute(1, 0, 2)
