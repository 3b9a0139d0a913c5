"""Marker -> body regressor training - host-side mirror of the reference's GAMMARegressorTrainOP
(motion/models/models_GAMMA_primitive.py:594-710: build_model :595-613, calc_loss :617-633, train :636-710). The
regressor forward over its recurrences, the 6-D -> axis-angle conversion, the SMPL-X marker loss, the full backward
(through SMPL-X skinning, the kinematic chain and the Gram-Schmidt rotation construction) and Adam run in the CUDA library
(eg_regressor_loss_backward / eg_adam_step_flat). Checkpoints keep the reference layout
{'epoch', 'model_state_dict', 'optimizer_state_dict'} -> <save_dir>/epoch-N.ckp.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import time

import numpy as np
import torch

from . import _lib, assets
from .models_gamma_primitive import REGRESSOR_CFG, MoshRegressor
from .smplx_parser import get_lbs_model

DEFAULT_LOSSCFG = {"weight_reg_hpose": 0.01}
DEFAULT_TRAINCFG = {"learning_rate": 3e-4, "batch_size": 64, "num_epochs": 100, "num_epochs_fix": 20, "saving_per_X_ep": 20,
                    "resume_training": False, "verbose": False, "save_dir": "results/checkpoints"}


class SyntheticBodyMarkerBatchGen:
    """Stand-in for BatchGeneratorAMASSCanonicalized.next_batch_genderselection (batch_gen_amass.py:348-429; licensed
    AMASS data absent): marker sequences produced by the surrogate body model from smooth random poses, so the loss has a
    reachable minimum. Yields (betas [B,T,10], markers [B,T,201])."""

    def __init__(self, lbs, n_seq, n_frames, device, seed=0):
        g = torch.Generator().manual_seed(seed)
        n = n_seq * n_frames
        xb = torch.zeros(n_seq, n_frames, 93)
        xb[..., :3] = torch.randn(n_seq, 1, 3, generator=g) * 0.3 + torch.cumsum(torch.randn(n_seq, n_frames, 3, generator=g) * 0.01, 1)
        xb[..., 3:69] = torch.randn(n_seq, 1, 66, generator=g) * 0.2 + torch.cumsum(torch.randn(n_seq, n_frames, 66, generator=g) * 0.01, 1)
        betas = (torch.randn(n_seq, 1, 10, generator=g) * 0.5).expand(n_seq, n_frames, 10).contiguous()
        _, _, mk = lbs.forward(xb.reshape(n, 93).to(device), betas.reshape(n, 10).to(device), want_joints=False, want_markers=True)
        self.markers = mk.reshape(n_seq, n_frames, -1).contiguous()
        self.betas = betas.to(device)
        self.index_rec = 0

    def has_next_rec(self):
        return self.index_rec < self.markers.shape[0]

    def reset(self):
        self.index_rec = 0

    def next_batch_genderselection(self, batch_size=64, gender="male", batch_first=True, noise=None):
        s = slice(self.index_rec, self.index_rec + batch_size)
        self.index_rec += batch_size
        return self.betas[s], self.markers[s]


class GAMMARegressorTrainOP:
    def __init__(self, modelconfig=None, lossconfig=None, trainconfig=None, device="cuda:0"):
        self.modelconfig = dict(modelconfig or REGRESSOR_CFG)
        self.lossconfig = dict(DEFAULT_LOSSCFG, **(lossconfig or {}))
        self.trainconfig = dict(DEFAULT_TRAINCFG, **(trainconfig or {}))
        self.device = torch.device(device)
        self.model = None
        self._h = None

    # ---- model + flat buffers -------------------------------------------------------------
    def build_model(self, seed=None):
        if seed is not None:
            torch.manual_seed(seed)
        self.model = MoshRegressor(self.modelconfig).to(self.device).train()
        self.use_cont = True
        if self.modelconfig.get("body_repr", "ssm2_67") != "ssm2_67":
            raise ValueError("other marker placement is not considered yet.")
        self.markers = self.model.markers = list(assets.marker_ids())
        self.bm = get_lbs_model(self.modelconfig.get("gender", "male"), self.device, marker_vids=assets.marker_ids())
        m = self.model
        ps = list(m.parameters())
        n = sum(p.numel() for p in ps)
        self.dims = _lib.EgRegressorDims(m.in_dim, m.h_dim, m.n_blocks, m.n_recur, m.body_dim)
        expect = _lib.lib().eg_regressor_param_count(C.byref(self.dims))
        if expect != n:
            raise _lib.EgError(f"parameter count {n} does not match the library layout {expect}")
        dev = self.device
        self.flat_params = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            k = p.numel()
            self.flat_params[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + k].view_as(p)
            p.grad = self.flat_grads[off:off + k].view_as(p)
            off += k
        h = C.c_void_p()
        _lib.check(_lib.lib().eg_regressor_train_create(C.byref(self.dims), _lib.ptr(self.flat_params), _lib.ptr(self.flat_grads),
                                                        self.bm._h, dev.index or 0, C.byref(h)))
        self._h = h
        self._stats = torch.zeros(3, dtype=torch.float32, device=dev)
        self._step = 0
        return self.model

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.lib().eg_regressor_train_destroy(self._h)
        except Exception:
            pass

    # ---- forward + loss + backward ----------------------------------------------------------
    def forward_loss_backward(self, marker_ref, betas):
        """xb_new = model(marker_ref, betas); optimizer.zero_grad(); calc_loss(...); loss.backward()  (:675-679).
        marker_ref [M,67,3] or [M,201], betas [M,10]. Returns (xb_new [M,93], loss, [loss_marker, loss_hpose]);
        the gradients are left in the flat gradient buffer (p.grad views)."""
        mk = marker_ref.reshape(marker_ref.shape[0], -1).to(torch.float32).contiguous()
        be = betas.reshape(-1, 10).to(torch.float32).contiguous()
        M = mk.shape[0]
        if mk.shape[1] != self.model.in_dim or be.shape[0] != M:
            raise _lib.EgError("marker_ref must be [M,201] and betas [M,10]")
        xb = torch.empty(M, 93, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_regressor_loss_backward(self._h, _lib.ptr(mk), _lib.ptr(be), M,
                                                             float(self.lossconfig["weight_reg_hpose"]), _lib.ptr(xb),
                                                             _lib.ptr(self._stats), _lib.stream_ptr(self.device)))
        s = self._stats.cpu().numpy()
        return xb, float(s[0]), np.array([s[1], s[2]])

    def optimizer_step(self, lr, betas=(0.9, 0.999), eps=1e-8):
        self._step += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_adam_step_flat(_lib.ptr(self.flat_params), _lib.ptr(self.flat_grads),
                                                    _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.flat_params.numel(),
                                                    float(lr), betas[0], betas[1], eps, 0.0, self._step,
                                                    _lib.stream_ptr(self.device)))

    def lr_at(self, epoch):
        tc = self.trainconfig                                     # get_scheduler 'lambda' rule (baseops.py:52-60)
        return tc["learning_rate"] * (1.0 - max(0, epoch - tc["num_epochs_fix"]) / float(tc["num_epochs"] - tc["num_epochs_fix"] + 1))

    def optimizer_state_dict(self):
        st, off = {}, 0
        for i, p in enumerate(self.model.parameters()):
            k = p.numel()
            st[i] = {"step": torch.tensor(float(self._step)), "exp_avg": self.exp_avg[off:off + k].view_as(p).clone(),
                     "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p).clone()}
            off += k
        return {"state": st, "param_groups": [{"lr": self.trainconfig["learning_rate"], "betas": (0.9, 0.999), "eps": 1e-8,
                                               "weight_decay": 0, "params": list(range(len(st)))}]}

    def load_optimizer_state_dict(self, sd):
        """optimizer.load_state_dict(checkpoint['optimizer_state_dict']) of the reference's resume path: the per-parameter
        Adam moments go back into the flat moment buffers and the step count (bias correction) continues."""
        off = 0
        for i, p in enumerate(self.model.parameters()):
            k = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1).to(self.exp_avg))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1).to(self.exp_avg_sq))
                self._step = int(float(st["step"]))
            off += k

    def train(self, batch_gen, log=print):
        """train (:636-710)."""
        self.build_model()
        tc = self.trainconfig
        start = 0
        if tc.get("resume_training"):
            ck = sorted(glob.glob(os.path.join(tc["save_dir"], "epoch-*.ckp")), key=os.path.getmtime)
            if ck:
                c = torch.load(ck[-1], map_location=self.device)
                self.model.load_state_dict(c["model_state_dict"])
                self.load_optimizer_state_dict(c["optimizer_state_dict"])
                start = c["epoch"]
            else:
                log("[INFO] resume_training set but no checkpoint under %s: training from scratch (as the reference does)" % tc["save_dir"])
        history = []
        for epoch in range(start, tc["num_epochs"]):
            tot, n, t0 = np.zeros(2), 0, time.time()
            lr = self.lr_at(epoch)
            while batch_gen.has_next_rec():
                data = batch_gen.next_batch_genderselection(tc["batch_size"], self.modelconfig.get("gender", "male"))
                if data is None:
                    continue
                batch_betas, marker_ref = data[:2]
                marker_ref = marker_ref.contiguous().view(-1, self.model.in_dim)
                batch_betas = batch_betas.contiguous().view(-1, 10)
                _, _, items = self.forward_loss_backward(marker_ref, batch_betas)
                self.optimizer_step(lr)
                tot += items; n += 1
            batch_gen.reset()
            tot /= max(n, 1)
            history.append(tot.copy())
            log("[epoch {:d}]:MSE_MARKER={:f}, MSE_HPOSE={:f}, time={:f}, lr={:f}".format(epoch + 1, tot[0], tot[1],
                                                                                        time.time() - t0, lr))
            if (1 + epoch) % tc["saving_per_X_ep"] == 0:
                os.makedirs(tc["save_dir"], exist_ok=True)
                torch.save({"epoch": epoch + 1, "model_state_dict": self.model.state_dict(),
                            "optimizer_state_dict": self.optimizer_state_dict()},
                           os.path.join(tc["save_dir"], "epoch-" + str(epoch + 1) + ".ckp"))
        return history


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu_index", type=int, default=0)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--num_epochs", type=int, default=2)
    ap.add_argument("--n_seq", type=int, default=512)
    ap.add_argument("--seq_len", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", a.gpu_index)
    op = GAMMARegressorTrainOP(trainconfig={"batch_size": a.batch_size, "num_epochs": a.num_epochs, "num_epochs_fix": 1,
                                            "saving_per_X_ep": 10 ** 9}, device=dev)
    lbs = get_lbs_model("male", dev, marker_vids=assets.marker_ids())
    op.train(SyntheticBodyMarkerBatchGen(lbs, a.n_seq, a.seq_len, dev))
