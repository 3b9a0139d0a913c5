"""Golden vectors for the start-body sampler from the REFERENCE'S OWN CrowdMotion.gen_init_body
(motion/exp_GAMMAPrimitive/utils/environments.py:1041-1131) - build container only. The method runs unmodified on the
reference's motion seed data/locomotion/subseq_00343.npz; absent third-party code is supplied by:
  self.bm_male(...)        -> the oracle's SMPL-X restatement on the surrogate body model
  pytorch3d.transforms     -> scipy.spatial.transform.Rotation (axis_angle_to_matrix, matrix_to_axis_angle,
                              euler_angles_to_matrix "XYZ" = Rx Ry Rz)
and torch.cuda.FloatTensor / .cuda() are aliased to their CPU forms. The two random draws of the method (start frame, yaw
jitter) are replayed from the same torch seed and stored next to the outputs.

Run:  python tests/golden/gen_sampler_golden.py   ->  tests/golden/sampler_golden.npz
"""
import os
import sys
import types

import numpy as np
import torch
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")

from egogen_b200 import assets                                # noqa: E402
from oracle import tgm as oracle_tgm                          # noqa: E402
from oracle.smplx_lbs import SMPLXOracle                      # noqa: E402


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
stub("pytorch3d")
stub("pytorch3d.structures")
stub("pytorch3d.transforms",
     axis_angle_to_matrix=lambda aa: f32(Rotation.from_rotvec(aa.double().numpy().reshape(-1, 3)).as_matrix()).reshape(aa.shape[:-1] + (3, 3)),
     matrix_to_axis_angle=lambda m: f32(Rotation.from_matrix(m.double().numpy().reshape(-1, 3, 3)).as_rotvec()).reshape(m.shape[:-2] + (3,)),
     euler_angles_to_matrix=lambda ang, convention: f32(Rotation.from_euler(convention, ang.double().numpy()).as_matrix()))
sys.modules["pytorch3d"].transforms = sys.modules["pytorch3d.transforms"]
sys.modules["pytorch3d"].structures = sys.modules["pytorch3d.structures"]
for n in ["smplx", "trimesh", "pyrender", "tensorboardX", "matplotlib", "matplotlib.pyplot", "omegaconf", "human_body_prior",
          "human_body_prior.tools", "shapely.plotting"]:
    stub(n)
stub("human_body_prior.tools.model_loader", load_vposer=None)
stub("exp_GAMMAPrimitive.utils.utils_canonicalize_babel", get_body_model=None, marker_ssm_67=list(range(67)))
stub("shapely", LineString=None, union_all=None, is_valid=None)
stub("shapely.geometry", MultiPoint=None, Point=None, Polygon=None, MultiPolygon=None, mapping=None, LinearRing=None)
sys.modules["tensorboardX"].SummaryWriter = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["torchgeometry"] = oracle_tgm
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.FloatTensor = lambda *a, device=None: torch.FloatTensor(*a)
sys.path.insert(0, os.path.join(REF, "motion"))
os.chdir(os.path.join(REF, "motion"))                        # gen_init_body opens data/locomotion/subseq_00343.npz relatively
from exp_GAMMAPrimitive.utils import environments as ref_envs   # noqa: E402


class BodyModel:
    """smplx body model of batch size 2, as far as gen_init_body uses it: missing parameters are the model's zeros"""

    def __init__(self):
        self.m = SMPLXOracle(assets.make_surrogate_smplx(seed=0))

    def __call__(self, **kw):
        n = kw["betas"].shape[0]
        z = lambda d: torch.zeros(n, d)
        return self.m.forward(betas=kw["betas"], global_orient=kw.get("global_orient", z(3)), body_pose=kw.get("body_pose", z(63)),
                              left_hand_pose=z(12), right_hand_pose=z(12), transl=kw.get("transl", z(3)))


me = types.SimpleNamespace(bm_male=BodyModel(), bm_female=None)
n_frames = len(np.load("data/locomotion/subseq_00343.npz")["poses"])
cases = [((0.5, -1.0, 0.0), (2.5, 1.5, 0.0), 3), ((-2.0, 0.3, 0.0), (1.0, 0.4, 0.0), 4), ((1.2, 2.2, 0.0), (-1.5, -2.0, 0.0), 5)]
rec = {k: [] for k in ("start", "target", "start_frame", "yaw", "transl", "global_orient", "wpath", "betas", "body_pose")}
for start, target, seed in cases:
    torch.manual_seed(seed)
    sf = torch.randint(0, n_frames - 1, (1,)).item()
    yaw = (torch.FloatTensor(1).uniform_(-1, 1) * torch.pi * 2 * 0.2).item()
    torch.manual_seed(seed)
    d = ref_envs.CrowdMotion.gen_init_body(me, np.array(start, dtype=np.float32), np.array(target, dtype=np.float32), True)
    ms = d["motion_seed"]
    for k, v in (("start", start), ("target", target), ("start_frame", sf), ("yaw", yaw), ("transl", ms["transl"]),
                 ("global_orient", ms["global_orient"]), ("wpath", d["wpath"]), ("betas", d["betas"]), ("body_pose", ms["body_pose"])):
        rec[k].append(np.asarray(v, dtype=np.float64 if k == "yaw" else np.float32))
    assert d["gender"] == "male" and d["floor_height"] == 0
np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), **{k: np.stack(v) for k, v in rec.items()})
print("wrote sampler_golden.npz", {k: np.stack(v).shape for k, v in rec.items()}, "frames", rec["start_frame"], "yaw", rec["yaw"])
