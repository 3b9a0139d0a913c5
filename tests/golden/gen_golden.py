"""Generate golden input/output vectors from the REFERENCE'S OWN classes (build container only).

The reference (ligengen/EgoGen, /root/reference) is imported with sys.modules stubs for the
third-party packages that are absent (smplx, torchgeometry, tensorboardX, matplotlib, omegaconf);
only code that does not touch those stubs is executed:
  crowd_ppo/utils.py:calc_sdf                       -> sdf_golden.npz
  models/models_GAMMA_primitive.py GAMMAPrimitiveVAE.sample_prior, MoshRegressor (use_cont=False
      and the 6-D pnet body via _forward), RotConverter.cont2rotmat  -> nets_motion_golden.npz
  models/models_policy_ppo.py GAMMAPolicyBase/Actor/Critic            -> nets_policy_golden.npz
  models/baseops.py CanonicalCoordinateExtractor                      -> coord_golden.npz
Weights are filled by egogen_b200.assets.fill_params_ (keyed by state_dict name) so the fixtures
carry only inputs and outputs. The committed .npz files travel to the GPU box; this script does not.

Run:  python tests/golden/gen_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")

for m in ["smplx", "torchgeometry", "tensorboardX", "matplotlib", "matplotlib.pyplot", "omegaconf"]:
    sys.modules[m] = types.ModuleType(m)
sys.modules["tensorboardX"].SummaryWriter = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, os.path.join(REF, "motion"))

from models import baseops as ref_baseops                    # noqa: E402
from models import models_policy_ppo as ref_policy           # noqa: E402
from models import models_GAMMA_primitive as ref_gamma       # noqa: E402
from crowd_ppo import utils as ref_utils                     # noqa: E402
from egogen_b200.assets import fill_params_                  # noqa: E402

torch.set_num_threads(4)
torch.manual_seed(1234)
g = torch.Generator().manual_seed(99)


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


# ---- calc_sdf -------------------------------------------------------------------------------
def sdf_case(D, npts, seed):
    gg = torch.Generator().manual_seed(seed)
    grid = torch.randn(D, generator=gg)
    center = torch.tensor([0.3, -0.2, 1.1])
    scale = torch.tensor([0.27])
    # in-range, out-of-range (border clamp) and exactly-on-cell-boundary points
    pts = (torch.rand(2, npts, 3, generator=gg) * 2 - 1) * 4.6 + center
    lin = ((2 * torch.arange(D[0]) + 1.0) / D[0] - 1.0) / scale + center[0]
    pts[0, :D[0], 0] = lin
    out = ref_utils.calc_sdf(pts.clone(), {"center": center, "scale": scale, "sdf": grid})
    return grid, center, scale, pts, out


g1 = sdf_case((7, 9, 11), 1500, 1)
g2 = sdf_case((32, 32, 32), 4000, 2)
save("sdf_golden.npz",
     grid_a=g1[0], center_a=g1[1], scale_a=g1[2], pts_a=g1[3], out_a=g1[4],
     grid_b=g2[0], center_b=g2[1], scale_b=g2[2], pts_b=g2[3], out_b=g2[4])

# ---- C-VAE predictor + regressor ------------------------------------------------------------
pred = ref_gamma.GAMMAPrimitiveVAE({"body_repr": "ssm2_67", "h_dim": 256, "z_dim": 128, "use_drnn_mlp": True,
                                    "hdims_mlp": [512, 256], "residual": True}).eval()
fill_params_(pred, seed=11)
reg_cont = ref_gamma.MoshRegressor({"body_repr": "ssm2_67", "h_dim": 128, "n_blocks": 10, "n_recur": 3,
                                    "actfun": "relu", "use_cont": True}).eval()
fill_params_(reg_cont, seed=12, w_gain=0.7)
B = 6
X = torch.randn(2, B, 201, generator=g) * 0.3
z = torch.randn(B, 128, generator=g)
betas = torch.randn(18 * B, 10, generator=g)
with torch.no_grad():
    Y = pred.sample_prior(X, z)                                             # [18,B,201]
    zero = lambda d: torch.zeros(18 * B, d)
    xb_cont = reg_cont._forward(Y.reshape(18 * B, -1), zero(3), zero(6), zero(126), zero(12), zero(12), betas)
    rot = ref_baseops.RotConverter.cont2rotmat(xb_cont[:, 3:135].contiguous().view(18 * B, -1, 6))
save("nets_motion_golden.npz", X=X, z=z, betas=betas, Y=Y, xb_cont=xb_cont, rotmat=rot)

# ---- policy nets ----------------------------------------------------------------------------
cfg = {"h_dim": 512, "z_dim": 128, "n_blocks": 2, "actfun": "lrelu", "body_repr": "ssm2_67_condi_marker_map",
       "min_logvar": -2.5, "max_logvar": 2.5}
actor, critic, shared = ref_policy.GAMMAActor(cfg), ref_policy.GAMMACritic(cfg), ref_policy.GAMMAPolicyBase(cfg)
fill_params_(actor, seed=21); fill_params_(critic, seed=22); fill_params_(shared, seed=23)
Bp = 8
obs = {"state": torch.randn(Bp, 2, 402, generator=g) * 0.5,
       "egosensing": torch.rand(Bp, 2, 32, generator=g) * 2 - 1,
       "dist": torch.rand(Bp, 1, generator=g), "time": 1 - torch.randint(0, 13, (Bp, 1), generator=g) / 13.0}
with torch.no_grad():
    hx = shared(obs)
    (mu, logvar), _ = actor(hx)
    val = critic(hx)
save("nets_policy_golden.npz", state=obs["state"], egosensing=obs["egosensing"], dist=obs["dist"],
     time=obs["time"], hx=hx, mu=mu, logvar=logvar, value=val)

# ---- canonical coordinate extractor ---------------------------------------------------------
ce = ref_baseops.CanonicalCoordinateExtractor(torch.device("cpu"))
jts = torch.randn(5, 22, 3, generator=g)
R, T = ce.get_new_coordinate_torch(jts.clone())
sub = np.load(os.path.join(REF, "motion/data/locomotion/subseq_00343.npz"))
jl = torch.as_tensor(sub["joints"][:1], dtype=torch.float32)
R2, T2 = ce.get_new_coordinate_torch(jl.clone())
save("coord_golden.npz", jts=jts, R=R, T=T, jts_loco=jl, R_loco=R2, T_loco=T2)


def gen_locomotion_seed():
    """tests/golden/locomotion_seed_00343.npz: the reference's motion seed fixture motion/data/locomotion/subseq_00343.npz
    trimmed to the keys the start-body samplers read (environments.py:1048-1066)."""
    d = np.load(os.path.join(REF, "motion", "data", "locomotion", "subseq_00343.npz"), allow_pickle=True)
    np.savez_compressed(os.path.join(HERE, "locomotion_seed_00343.npz"), poses=d["poses"][:, :66].astype(np.float32),
                        trans=d["trans"].astype(np.float32), betas=d["betas"].astype(np.float32))


gen_locomotion_seed()
