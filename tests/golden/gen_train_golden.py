"""Golden vectors for the motion-model TRAINING objectives from the REFERENCE'S OWN train ops (build container only).

motion/models/models_GAMMA_primitive.py is imported as in gen_golden.py; the three train ops are created without their
__init__ (which opens log directories / tensorboard writers and loads SMPL-X files) and given the attributes it would
set. The two third-party pieces their losses call are supplied by the oracle's restatements, so what is pinned here is
the reference's own objective code (loss terms, weights, tensor slicing, which parameters receive gradients):
  torchgeometry  -> oracle.tgm        (RotConverter.cont2aa inside MoshRegressor.forward)
  self.bm(...)   -> oracle.smplx_lbs  (SMPL-X vertices / joints of the surrogate body model, seed 0)
The reparameterisation noise is torch.randn_like under a fixed seed; the same draw is stored as `eps`.

  GAMMAPrimitiveVAETrainOP.calc_loss / calc_loss_rollout      (:413-503)
  GAMMARegressorTrainOP: model(marker_ref, betas) + calc_loss (:617-633, :675-679)
  GAMMAPrimitiveComboTrainOP.calc_loss_one                    (:787-838), with and without scheduled sampling

Run:  python tests/golden/gen_train_golden.py   ->  tests/golden/train_golden.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")

from oracle import tgm as oracle_tgm                          # noqa: E402
from oracle.smplx_lbs import SMPLXOracle                      # noqa: E402
from egogen_b200 import assets                                # noqa: E402
from egogen_b200.assets import fill_params_                   # noqa: E402

for m in ["smplx", "tensorboardX", "matplotlib", "matplotlib.pyplot", "omegaconf"]:
    sys.modules[m] = types.ModuleType(m)
sys.modules["torchgeometry"] = oracle_tgm
sys.modules["tensorboardX"].SummaryWriter = object
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, os.path.join(REF, "motion"))
from models import models_GAMMA_primitive as ref_gamma       # noqa: E402

torch.set_num_threads(4)
PRED_CFG = {"body_repr": "ssm2_67", "h_dim": 256, "z_dim": 128, "t_his": 2, "t_pred": 18, "use_drnn_mlp": True,
            "hdims_mlp": [512, 256], "residual": True}
REG_CFG = {"gender": "male", "h_dim": 128, "n_blocks": 10, "n_recur": 3, "body_repr": "ssm2_67", "actfun": "relu",
           "use_cont": True}
LOSS = {"weight_rec": 1.0, "weight_td": 3.0, "weight_kld": 1.0, "annealing_kld": False, "robust_kld": True,
        "weight_reg_hpose": 0.01}
MARKERS = assets.marker_ids()
g = torch.Generator().manual_seed(2024)
out = {}


class BodyModel:
    """stands where smplx.create(...) does: bm(return_verts=True, **body_param) -> .vertices, .joints"""

    def __init__(self):
        self.m = SMPLXOracle(assets.make_surrogate_smplx(seed=0))

    def __call__(self, return_verts=True, **kw):
        return self.m.forward(betas=kw["betas"], global_orient=kw["global_orient"], body_pose=kw["body_pose"],
                              left_hand_pose=kw["left_hand_pose"], right_hand_pose=kw["right_hand_pose"], transl=kw["transl"])


def grad_norms(module):
    return np.array([p.grad.norm().item() if p.grad is not None else -1.0 for p in module.parameters()], dtype=np.float64)


def regressor_weights(reg, seed):
    fill_params_(reg, seed=seed, w_gain=0.5)
    with torch.no_grad():        # away from the degenerate all-zero 6-D rotation
        gg = torch.Generator().manual_seed(seed)
        reg.pnet.out_fc.bias[3:135] = torch.tensor([1.0, 0.0, 0.0, 1.0, 0.0, 0.0]).repeat(22) + torch.randn(132, generator=gg) * 0.3


# ---- predictor: calc_loss and calc_loss_rollout --------------------------------------------------
op = object.__new__(ref_gamma.GAMMAPrimitiveVAETrainOP)
op.modelconfig, op.lossconfig, op.trainconfig, op.device = PRED_CFG, LOSS, {"num_epochs": 400, "max_rollout": 8}, torch.device("cpu")
op.build_model()
fill_params_(op.model, seed=41)
B = 5
data = torch.cumsum(torch.randn(20, B, 201, generator=g) * 0.02, dim=0) + torch.randn(1, B, 201, generator=g) * 0.3
torch.manual_seed(77); eps = torch.randn(B, 128)
torch.manual_seed(77)
loss, info = op.calc_loss(data, 0)
loss.backward()
out.update(p_data=data, p_eps=eps, p_loss=loss.item(), p_info=info, p_gradnorm=grad_norms(op.model),
           p_grad_dout_bias=op.model.d_out.bias.grad.clone())
op.model.zero_grad()
n_t = 61
walk = torch.cumsum(torch.randn(n_t, B, 1, 3, generator=g) * 0.01, dim=0)
mk = (torch.randn(1, B, 67, 3, generator=g) * 0.3 + walk + torch.cumsum(torch.randn(n_t, B, 67, 3, generator=g) * 0.002, dim=0)).reshape(n_t, B, 201)
j0 = torch.randn(1, B, 22, 3, generator=g) * 0.3
j0[:, :, 1, 0] += 0.5; j0[:, :, 2, 0] -= 0.5
jts = (j0 + walk).contiguous()
torch.manual_seed(78); eps_r = torch.stack([torch.randn(B, 128) for _ in range(3)])
torch.manual_seed(78)
loss, info = op.calc_loss_rollout((mk, jts.reshape(n_t, B, -1)), 0)
loss.backward()
out.update(r_markers=mk, r_jts=jts, r_eps=eps_r, r_loss=loss.item(), r_info=info, r_gradnorm=grad_norms(op.model))

# ---- regressor: forward + calc_loss ----------------------------------------------------------------
op2 = object.__new__(ref_gamma.GAMMARegressorTrainOP)
op2.modelconfig, op2.lossconfig, op2.trainconfig, op2.device = REG_CFG, LOSS, {}, torch.device("cpu")
op2.model = ref_gamma.MoshRegressor(REG_CFG).train()
op2.model.markers = op2.markers = MARKERS
op2.bm = BodyModel()
regressor_weights(op2.model, 3)
M = 9
xb_true = torch.randn(M, 93, generator=g) * 0.3
betas = torch.randn(M, 10, generator=g) * 0.5
with torch.no_grad():
    kw = dict(transl=xb_true[:, :3], global_orient=xb_true[:, 3:6], body_pose=xb_true[:, 6:69], left_hand_pose=xb_true[:, 69:81],
              right_hand_pose=xb_true[:, 81:], betas=betas)
    marker_ref = op2.bm(**kw).vertices[:, MARKERS, :] + torch.randn(M, 67, 3, generator=g) * 0.01
xb_new = op2.model(marker_ref.detach(), betas)
loss, items = op2.calc_loss(marker_ref, xb_new, betas)
loss.backward()
out.update(g_marker_ref=marker_ref, g_betas=betas, g_xb=xb_new.detach(), g_loss=loss.item(), g_items=items,
           g_gradnorm=grad_norms(op2.model), g_grad_out_bias=op2.model.pnet.out_fc.bias.grad.clone())

# ---- combo: calc_loss_one ------------------------------------------------------------------------------
for sched in (0, 1):
    op3 = object.__new__(ref_gamma.GAMMAPrimitiveComboTrainOP)
    op3.lossconfig, op3.trainconfig = LOSS, {"batch_size": B, "num_epochs": 400}
    op3.t_his, op3.t_pred, op3.n_meshes, op3.use_scheduled_sampling = 2, 18, B * 18, bool(sched)
    op3.model = ref_gamma.GAMMAPrimitiveCombo(PRED_CFG, REG_CFG)
    op3.model.predictor.train()
    fill_params_(op3.model.predictor, seed=41)
    regressor_weights(op3.model.regressor, 5)
    op3.bm, op3.markers = BodyModel(), MARKERS
    cb = (torch.randn(1, B, 10, generator=torch.Generator().manual_seed(6)) * 0.5).expand(20, B, 10).contiguous()
    torch.manual_seed(79); eps_c = torch.randn(B, 128)
    torch.manual_seed(79)
    loss, info = op3.calc_loss_one([cb, data], 0)
    loss.backward()
    out.update({f"c{sched}_loss": loss.item(), f"c{sched}_info": info, f"c{sched}_gradnorm": grad_norms(op3.model.predictor),
                f"c{sched}_reg_has_grad": np.array(int(any(p.grad is not None and p.grad.abs().sum() > 0 for p in op3.model.regressor.parameters())))})
    out.update(c_betas=cb, c_eps=eps_c)

# ---- learning-rate rule of the train loops (baseops.get_scheduler 'lambda', stepped once per epoch :526-529,560) ------
from models import baseops as ref_baseops                     # noqa: E402
w = torch.nn.Parameter(torch.zeros(1))
opt = torch.optim.Adam([w], lr=5e-4)
sched = ref_baseops.get_scheduler(opt, policy="lambda", num_epochs_fix=100, num_epochs=400)
lrs = []
for epoch in range(400):
    lrs.append(opt.param_groups[0]["lr"])                    # the rate the epoch's optimiser steps use
    opt.step(); sched.step()
out["sched_lr"] = np.array(lrs, dtype=np.float64)

np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **{k: np.asarray(v) for k, v in out.items()})
print("wrote train_golden.npz", {k: np.asarray(v).shape for k, v in out.items()})
