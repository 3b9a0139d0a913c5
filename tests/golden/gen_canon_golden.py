"""Golden primitive from the REFERENCE'S OWN canonicalize_subsequence
(motion/exp_GAMMAPrimitive/utils/utils_canonicalize_samp.py:123-187) - build container only. The function (with the
module's get_new_coordinate / calc_calibrate_offset) runs unmodified on a synthetic SAMP-format recording; smplx.create is
served by the oracle's SMPL-X restatement on the surrogate body model, `.cuda()` is aliased away.

Run:  python tests/golden/gen_canon_golden.py   ->  tests/golden/canon_golden.npz
"""
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")

from egogen_b200 import assets                                # noqa: E402
from oracle.smplx_lbs import SMPLXOracle                      # noqa: E402


class BodyModel:
    def __init__(self):
        self.m = SMPLXOracle(assets.make_surrogate_smplx(seed=0))

    def cuda(self):
        return self

    def __call__(self, return_verts=True, **kw):
        n = kw["betas"].shape[0]
        z = lambda d: torch.zeros(n, d)
        return self.m.forward(betas=kw["betas"], global_orient=kw.get("global_orient", z(3)), body_pose=kw.get("body_pose", z(63)),
                              left_hand_pose=z(12), right_hand_pose=z(12), transl=kw.get("transl", z(3)))


smplx_stub = types.ModuleType("smplx")
smplx_stub.create = lambda *a, **k: BodyModel()
sys.modules["smplx"] = smplx_stub
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, os.path.join(REF, "motion"))
os.chdir(os.path.join(REF, "motion"))                        # the module opens data/CMU.json and data/SSM2.json relatively
from exp_GAMMAPrimitive.utils import utils_canonicalize_samp as ref_canon   # noqa: E402


def synthetic_recording(seed, T):
    g = torch.Generator().manual_seed(seed)
    transl = torch.randn(1, 3, generator=g) * 0.5 + torch.cumsum(torch.randn(T, 3, generator=g) * 0.01, 0)
    pose = torch.randn(1, 165, generator=g) * 0.25 + torch.cumsum(torch.randn(T, 165, generator=g) * 0.005, 0)
    pose[:, :3] += torch.tensor([1.2, 0.3, -0.4])
    return transl.numpy(), pose.numpy(), (torch.randn(16, generator=g) * 0.5).numpy()


transl, pose, betas = synthetic_recording(4, 200)
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "locomotion_synth.pkl")
    with open(path, "wb") as f:
        pickle.dump({"mocap_framerate": 120.0, "pose_est_trans": transl, "pose_est_fullposes": pose, "shape_est_betas": betas}, f)
    assert ref_canon.canonicalize_subsequence(path, 150, 210) is None
    out = ref_canon.canonicalize_subsequence(path, 30, 90)
assert out["gender"] == "male" and out["mocap_framerate"] == 120
arrs = {k: np.asarray(v) for k, v in out.items() if k not in ("gender", "mocap_framerate")}
arrs.update(in_transl=transl, in_pose=pose, in_betas=betas)
np.savez_compressed(os.path.join(HERE, "canon_golden.npz"), **arrs)
print("wrote canon_golden.npz", {k: v.shape for k, v in arrs.items()})
