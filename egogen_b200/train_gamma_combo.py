"""Joint predictor + regressor objective - host-side mirror of the reference's GAMMAPrimitiveComboTrainOP
(motion/models/models_GAMMA_primitive.py:713-1093): build_model :735-770, calc_loss_marker :797-815,
calc_loss_regressor :787-794, calc_loss_one :819-838, train :1017-1093 (Adam over the PREDICTOR's parameters only; the
regressor is a fixed differentiable layer between the predicted markers and the SMPL-X cycle loss).

One step = eg_cvae_forward_train -> eg_regressor_cycle_backward (regressor forward, 6-D -> axis-angle, SMPL-X markers, the
marker / smoothness / hand terms and their gradient w.r.t. the predicted markers) -> eg_cvae_backward (+ that gradient) ->
eg_adam_step_flat. The reference's two rollout variants (calc_loss_rollout[_UseGTTransform] :841-1014) hand 4-D joint
tensors [t,b,J,3] to CanonicalCoordinateExtractor.get_new_coordinate_torch, which indexes them as [b,J,3]
(baseops.py:214-225) and fails in torch.cross; they have no defined result to mirror and are not provided.
"""
from __future__ import annotations

import glob
import os
import time

import numpy as np
import torch

from . import _lib
from .models_gamma_primitive import PREDICTOR_CFG, REGRESSOR_CFG, GAMMAPrimitiveCombo
from .train_gamma_predictor import DEFAULT_LOSSCFG as _PRED_LOSS, GAMMAPrimitiveVAETrainOP
from .train_gamma_regressor import GAMMARegressorTrainOP

DEFAULT_LOSSCFG = dict(_PRED_LOSS, weight_reg_hpose=0.01)
DEFAULT_TRAINCFG = {"learning_rate": 1e-4, "batch_size": 32, "num_epochs": 100, "num_epochs_fix": 20, "saving_per_X_ep": 20,
                    "resume_training": False, "verbose": False, "save_dir": "results/checkpoints",
                    "scheduled_sampling": False}


class GAMMAPrimitiveComboTrainOP:
    def __init__(self, predictorcfg=None, regressorcfg=None, lossconfig=None, trainconfig=None, device="cuda:0"):
        self.predictorcfg = dict(predictorcfg or PREDICTOR_CFG)
        self.regressorcfg = dict(regressorcfg or REGRESSOR_CFG)
        self.lossconfig = dict(DEFAULT_LOSSCFG, **(lossconfig or {}))
        self.trainconfig = dict(DEFAULT_TRAINCFG, **(trainconfig or {}))
        self.device = torch.device(device)
        self.t_his = self.predictorcfg.get("t_his", 2)
        self.t_pred = self.predictorcfg.get("t_pred", 18)
        self.use_scheduled_sampling = self.trainconfig.get("scheduled_sampling", False)
        self.model = None

    def build_model(self, seed=None, predictor_state=None, regressor_state=None):
        """The combo model over two flat parameter buffers (predictor: trained; regressor: fixed). Pre-trained
        state_dicts (the reference loads epoch-300 / epoch-100 checkpoints, :758-769) can be passed in."""
        self._pop = GAMMAPrimitiveVAETrainOP(self.predictorcfg, self.lossconfig, self.trainconfig, device=self.device)
        self._rop = GAMMARegressorTrainOP(self.regressorcfg, self.lossconfig, self.trainconfig, device=self.device)
        self._pop.build_model(seed)
        self._rop.build_model(seed)
        if predictor_state is not None:
            self._pop.model.load_state_dict(predictor_state)
        if regressor_state is not None:
            self._rop.model.load_state_dict(regressor_state)
        self.model = GAMMAPrimitiveCombo(self.predictorcfg, self.regressorcfg)
        self.model.predictor, self.model.regressor = self._pop.model, self._rop.model     # the flat-buffer-backed modules
        self.model.predictor.train()
        self._stats = torch.zeros(4, dtype=torch.float32, device=self.device)       # loss, rec, kld, (unused)
        self._rstats = torch.zeros(2, dtype=torch.float32, device=self.device)      # reg marker term, hand term
        self.gen = None
        return self.model

    # ---- one primitive -----------------------------------------------------------------------
    def calc_loss_one(self, data, epoch, eps=None):
        """calc_loss_one (:819-838): data = [betas [20,B,10], ref_markers [20,B,>=201], ...] (time-major). Gradients of
        loss_marker + loss_bparams w.r.t. the predictor are left in its flat gradient buffer.
        Returns (loss, [REC, KLD, REG, HPOSE], Yb_rec [18,B,93])."""
        betas, ref = data[:2]
        X = ref[:self.t_his].to(torch.float32).contiguous()
        Y = ref[self.t_his:, :, :67 * 3].to(torch.float32).contiguous()
        betas_Y = betas[self.t_his:].to(torch.float32).contiguous()
        T, B = Y.shape[:2]
        if eps is None:
            eps = torch.randn(B, self._pop.model.z_dim, device=self.device, generator=self.gen)
        eps = eps.contiguous()
        lc = self.lossconfig
        w_kld = self._pop._weight_kld(epoch)
        Y_rec = torch.empty_like(Y)
        dY = torch.empty_like(Y)
        Yb = torch.empty(T * B, 93, device=self.device)
        self._stats.zero_(); self._rstats.zero_(); self._pop.flat_grads.zero_()
        L, st = _lib.lib(), _lib.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _lib.check(L.eg_cvae_forward_train(self._pop._h, _lib.ptr(X), _lib.ptr(Y), _lib.ptr(eps), B, _lib.ptr(Y_rec), st))
            _lib.check(L.eg_regressor_cycle_backward(self._rop._h, _lib.ptr(Y_rec), _lib.ptr(betas_Y), _lib.ptr(Y), T, B,
                                                     float(lc["weight_rec"]), float(lc["weight_td"]),
                                                     float(lc["weight_reg_hpose"]), 1.0, 1, _lib.ptr(dY), _lib.ptr(Yb),
                                                     _lib.ptr(self._rstats), st))
            _lib.check(L.eg_cvae_backward(self._pop._h, _lib.ptr(X), _lib.ptr(Y), _lib.ptr(eps), B, float(lc["weight_rec"]),
                                          float(lc["weight_td"]), w_kld, int(bool(lc["robust_kld"])), 1.0,
                                          int(bool(self.use_scheduled_sampling)), _lib.ptr(Y_rec), _lib.ptr(dY),
                                          _lib.ptr(self._stats), st))
        s, r = self._stats.cpu().numpy(), self._rstats.cpu().numpy()
        loss = float(s[0] + r[0] + lc["weight_reg_hpose"] * r[1])
        return loss, np.array([s[1], s[2], r[0], r[1]]), Yb.view(T, B, 93)

    def optimizer_step(self, lr):
        self._pop.optimizer_step(lr)                      # optim.Adam(self.model.predictor.parameters()) (:1024)

    def train(self, batch_gen, log=print):
        """train (:1017-1093) with calc_loss_one; batches come from next_batch_genderselection(batch_first=False)."""
        if self.model is None:
            self.build_model()
        tc = self.trainconfig
        start = 0
        if tc.get("resume_training"):
            ck = sorted(glob.glob(os.path.join(tc["save_dir"], "epoch-*.ckp")), key=os.path.getmtime)
            if ck:
                c = torch.load(ck[-1], map_location=self.device)
                self.model.load_state_dict(c["model_state_dict"])
                self._pop.load_optimizer_state_dict(c["optimizer_state_dict"])
                start = c["epoch"]
            else:
                log("[INFO] resume_training set but no checkpoint under %s: training from scratch (as the reference does)" % tc["save_dir"])
        history = []
        for epoch in range(start, tc["num_epochs"]):
            tot, n, t0 = np.zeros(4), 0, time.time()
            lr = self._pop.lr_at(epoch)
            while batch_gen.has_next_rec():
                data = batch_gen.next_batch_genderselection(batch_size=tc["batch_size"],
                                                            gender=self.regressorcfg.get("gender", "male"), batch_first=False)
                if data is None:
                    continue
                _, items, _ = self.calc_loss_one(data, epoch)
                self.optimizer_step(lr)
                tot += items; n += 1
            batch_gen.reset()
            tot /= max(n, 1)
            history.append(tot.copy())
            log("[epoch {:d}]:REC={:f}, KLD={:f}, REG={:f}, HPOSE={:f}, time={:f}, lr={:f}".format(
                epoch + 1, tot[0], tot[1], tot[2], tot[3], time.time() - t0, lr))
            if (1 + epoch) % tc["saving_per_X_ep"] == 0:
                os.makedirs(tc["save_dir"], exist_ok=True)
                torch.save({"epoch": epoch + 1, "model_state_dict": self.model.state_dict(),
                            "optimizer_state_dict": self._pop.optimizer_state_dict()},
                           os.path.join(tc["save_dir"], "epoch-" + str(epoch + 1) + ".ckp"))
        return history
