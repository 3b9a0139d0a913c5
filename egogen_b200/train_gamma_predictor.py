"""C-VAE marker-predictor training (BASELINE config 3) - host-side mirror of the reference's
GAMMAPrimitiveVAETrainOP (motion/models/models_GAMMA_primitive.py:389-589: build_model, calc_loss :413-432,
calc_loss_rollout :435-503, train :507-589) and of the entry script
motion/exp_GAMMAPrimitive/train_GAMMAPredictor.py:50-59. Forward, losses, backward (BPTT) and Adam run in the CUDA
library (eg_cvae_loss_backward / eg_adam_step_flat); the rollout's frame changes use eg_new_coordinate /
eg_rigid_points. Checkpoints keep the reference layout {'epoch', 'model_state_dict', 'optimizer_state_dict'} ->
<save_dir>/epoch-N.ckp.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import time

import numpy as np
import torch

from . import _lib
from .models_gamma_primitive import PREDICTOR_CFG, GAMMAPrimitiveVAE

DEFAULT_LOSSCFG = {"weight_rec": 1.0, "weight_td": 3.0, "weight_kld": 1.0, "annealing_kld": False, "robust_kld": True}
DEFAULT_TRAINCFG = {"max_rollout": 8, "learning_rate": 5e-4, "batch_size": 128, "num_epochs": 400, "num_epochs_fix": 100,
                    "saving_per_X_ep": 100, "resume_training": False, "verbose": False, "save_dir": "results/checkpoints"}


class SyntheticPrimitiveBatchGen:
    """Stand-in for BatchGeneratorAMASSCanonicalized (batch_gen_amass.py:61-429, licensed SAMP data absent): smooth
    random-walk marker sequences [N, T, 201] + joints [N, T, 22, 3] with the same iteration interface
    (has_next_rec / next_batch / next_batch_with_jts / reset / reset_with_jts)."""

    def __init__(self, n_seq, n_frames, device, seed=0):
        g = torch.Generator().manual_seed(seed)
        base = torch.randn(n_seq, 1, 67, 3, generator=g) * 0.3
        walk = torch.cumsum(torch.randn(n_seq, n_frames, 1, 3, generator=g) * 0.01, dim=1)
        jit = torch.cumsum(torch.randn(n_seq, n_frames, 67, 3, generator=g) * 0.002, dim=1)
        self.data_all = (base + walk + jit).reshape(n_seq, n_frames, 201).to(device)
        j = torch.randn(n_seq, 1, 22, 3, generator=g) * 0.3
        j[:, :, 1, 0] += 0.5; j[:, :, 2, 0] -= 0.5                   # hips apart so the canonical frame is defined
        self.jts_all = (j + walk).to(device).contiguous()
        self.index_rec = 0

    def has_next_rec(self):
        return self.index_rec < self.data_all.shape[0]

    def reset(self):
        self.index_rec = 0

    reset_with_jts = reset

    def next_batch(self, batch_size=64, noise=None):
        d = self.data_all[self.index_rec:self.index_rec + batch_size, :20].permute(1, 0, 2).contiguous()
        self.index_rec += batch_size
        return d

    def next_batch_with_jts(self, batch_size=64, noise=None):
        d = self.data_all[self.index_rec:self.index_rec + batch_size].permute(1, 0, 2)
        j = self.jts_all[self.index_rec:self.index_rec + batch_size].permute(1, 0, 2, 3)
        self.index_rec += batch_size
        return d, j


class GAMMAPrimitiveVAETrainOP:
    def __init__(self, modelconfig=None, lossconfig=None, trainconfig=None, device="cuda:0"):
        self.modelconfig = dict(modelconfig or PREDICTOR_CFG)
        self.lossconfig = dict(DEFAULT_LOSSCFG, **(lossconfig or {}))
        self.trainconfig = dict(DEFAULT_TRAINCFG, **(trainconfig or {}))
        self.device = torch.device(device)
        self.model = None
        self._h = None
        self.gen = None

    # ---- model + flat buffers -------------------------------------------------------------
    def build_model(self, seed=None):
        if seed is not None:
            torch.manual_seed(seed)
        self.model = GAMMAPrimitiveVAE(self.modelconfig).to(self.device).train()
        self.max_rollout = self.trainconfig.get("max_rollout", None)
        self.t_his = self.modelconfig.get("t_his", 2)
        ps = list(self.model.parameters())
        n = sum(p.numel() for p in ps)
        m = self.model
        self.dims = _lib.EgCvaeDims(m.in_dim, m.h_dim, m.z_dim, m.d_mlp.layers[0].out_features)
        expect = _lib.lib().eg_cvae_param_count(C.byref(self.dims))
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
        _lib.check(_lib.lib().eg_cvae_create(C.byref(self.dims), _lib.ptr(self.flat_params), _lib.ptr(self.flat_grads),
                                             dev.index or 0, C.byref(h)))
        self._h = h
        self._stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self._step = 0
        return self.model

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.lib().eg_cvae_destroy(self._h)
        except Exception:
            pass

    # ---- losses -----------------------------------------------------------------------------
    def _weight_kld(self, epoch):
        w = self.lossconfig["weight_kld"]
        if self.lossconfig["annealing_kld"]:
            w = min(float(epoch) / (0.9 * self.trainconfig["num_epochs"]), 1.0) * w
        return float(w)

    def _primitive(self, X, Y, epoch, scale, eps=None):
        """forward + loss + backward of one primitive; X [2,B,201], Y [18,B,201]. Returns Y_rec."""
        B = X.shape[1]
        X = X.to(torch.float32).contiguous(); Y = Y.to(torch.float32).contiguous()
        if eps is None:
            eps = torch.randn(B, self.model.z_dim, device=self.device, generator=self.gen)
        Y_rec = torch.empty_like(Y)
        lc = self.lossconfig
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_cvae_loss_backward(self._h, _lib.ptr(X), _lib.ptr(Y), _lib.ptr(eps.contiguous()), B,
                                                        float(lc["weight_rec"]), float(lc["weight_td"]),
                                                        self._weight_kld(epoch), int(bool(lc["robust_kld"])), float(scale),
                                                        _lib.ptr(Y_rec), _lib.ptr(self._stats), _lib.stream_ptr(self.device)))
        return Y_rec

    def calc_loss(self, data, epoch, eps=None):
        """calc_loss (:413-432): data [20,B,201]. Gradients are left in the flat gradient buffer.
        Returns (loss, [loss, rec, kld])."""
        self._stats.zero_(); self.flat_grads.zero_()
        self._primitive(data[:self.t_his], data[self.t_his:, :, :self.model.in_dim], epoch, 1.0, eps)
        s = self._stats.cpu().numpy()
        return float(s[0]), np.array([s[0], s[1], s[2]])

    def _frames(self, jts0):
        B = jts0.shape[0]
        j = jts0.reshape(B, -1).to(torch.float32).contiguous()
        R = torch.empty(B, 3, 3, device=self.device); T = torch.empty(B, 3, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_new_coordinate(_lib.ptr(j), j.shape[1], B, _lib.ptr(R), _lib.ptr(T),
                                                    _lib.stream_ptr(self.device)))
        return R, T

    def _rigid(self, R, T, pts, inverse):
        nt, B = pts.shape[:2]
        p = pts.reshape(nt, B, -1, 3).to(torch.float32).contiguous()
        out = torch.empty_like(p)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().eg_rigid_points(_lib.ptr(R.contiguous()), _lib.ptr(T.contiguous()), _lib.ptr(p), nt, B,
                                                  p.shape[2], int(inverse), _lib.ptr(out), _lib.stream_ptr(self.device)))
        return out

    def calc_loss_rollout(self, data, epoch, eps_list=None):
        """calc_loss_rollout (:435-503): chained 20-frame primitives; from the second primitive on the motion seed is the
        previous PREDICTION re-expressed in the new canonical frame (detached), the target the transformed ground truth."""
        ref_markers, ref_jts = data
        n_t, n_b = ref_markers.shape[:2]
        ref_jts = ref_jts.contiguous().view(n_t, n_b, -1, 3)
        t_his, t_pred = self.t_his, 20 - self.t_his
        # number of primitives first (the loss is their mean)
        n_prim, t = 0, 0
        while t < n_t and t + 20 < n_t and n_prim < self.max_rollout:
            n_prim += 1; t += t_pred
        self._stats.zero_(); self.flat_grads.zero_()
        t, k = 0, 0
        Y_rec = R_prev = T_prev = None
        while k < n_prim:
            mk = ref_markers[t:t + 20]; jt = ref_jts[t:t + 20]
            if t == 0:
                X = mk[:t_his]; Y = mk[t_his:, :, :self.model.in_dim]
                R_prev, T_prev = self._frames(jt[0])
            else:
                R_cur, T_cur = self._frames(jt[0])
                Y = self._rigid(R_cur, T_cur, mk[t_his:, :, :self.model.in_dim], True).reshape(t_pred, n_b, -1)
                Xg = self._rigid(R_prev, T_prev, Y_rec[-t_his:], False)
                X = self._rigid(R_cur, T_cur, Xg, True).reshape(t_his, n_b, -1)
                R_prev, T_prev = R_cur, T_cur
            Y_rec = self._primitive(X, Y, epoch, 1.0 / n_prim, None if eps_list is None else eps_list[k])
            t += t_pred; k += 1
        s = self._stats.cpu().numpy()
        return float(s[0]), np.array([s[0], s[1], s[2]])

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
        """train (:507-589)."""
        self.build_model()
        tc = self.trainconfig
        start = 0
        if tc.get("resume_training"):
            ck = sorted(glob.glob(os.path.join(tc["save_dir"], "epoch-*.ckp")), key=os.path.getmtime)
            if not ck:
                raise FileExistsError("the pre-trained checkpoint does not exist.")
            c = torch.load(ck[-1], map_location=self.device)
            self.model.load_state_dict(c["model_state_dict"])
            if not tc.get("fine_tune", False):
                self.load_optimizer_state_dict(c["optimizer_state_dict"])
                start = c["epoch"]
        for epoch in range(start, tc["num_epochs"]):
            tot, n, t0 = np.zeros(3), 0, time.time()
            lr = self.lr_at(epoch)
            while batch_gen.has_next_rec():
                if self.max_rollout is None:
                    data = batch_gen.next_batch(tc["batch_size"])
                    _, items = self.calc_loss(data, epoch)
                else:
                    data = batch_gen.next_batch_with_jts(tc["batch_size"])
                    _, items = self.calc_loss_rollout(data, epoch)
                self.optimizer_step(lr)
                tot += items; n += 1
            batch_gen.reset() if self.max_rollout is None else batch_gen.reset_with_jts()
            tot /= max(n, 1)
            log("[epoch {:d}]:ALL={:f}, REC={:f}, KLD={:f}, time={:f}, lr={:f}".format(epoch + 1, tot[0], tot[1], tot[2],
                                                                                      time.time() - t0, lr))
            if (1 + epoch) % tc["saving_per_X_ep"] == 0:
                os.makedirs(tc["save_dir"], exist_ok=True)
                torch.save({"epoch": epoch + 1, "model_state_dict": self.model.state_dict(),
                            "optimizer_state_dict": self.optimizer_state_dict()},
                           os.path.join(tc["save_dir"], "epoch-" + str(epoch + 1) + ".ckp"))


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="MPVAE_samp20_2frame_rollout")
    ap.add_argument("--gpu_index", type=int, default=0)
    ap.add_argument("--batch_size", type=int, default=128)
    ap.add_argument("--num_epochs", type=int, default=2)
    ap.add_argument("--n_seq", type=int, default=1024)
    a = ap.parse_args()
    dev = torch.device("cuda", a.gpu_index)
    op = GAMMAPrimitiveVAETrainOP(trainconfig={"batch_size": a.batch_size, "num_epochs": a.num_epochs, "num_epochs_fix": 1,
                                               "saving_per_X_ep": 10 ** 9}, device=dev)
    op.train(SyntheticPrimitiveBatchGen(a.n_seq, 200, dev))
