"""Motion-primitive networks - host-side mirror of the reference's
motion/models/models_GAMMA_primitive.py (GAMMAPrimitiveVAE :36-156, MoshRegressor :178-301,
GAMMAPrimitiveCombo :307-386, GAMMAPrimitiveComboGenOP :1099-1139).

The modules below only HOLD parameters, under the reference's state_dict names and shapes, so the
reference's ``epoch-*.ckp`` checkpoints load unchanged; the arithmetic of ``sample_prior`` runs in the
CUDA library (eg_motion_sample_prior). There is no torch forward and no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import nn

from . import _lib, assets


class _MLP(nn.Module):
    """Parameter container with the layout of baseops.MLP (layers.{i}.{weight,bias})."""

    def __init__(self, in_dim, h_dims):
        super().__init__()
        self.layers = nn.ModuleList()
        d = in_dim
        for h in h_dims:
            self.layers.append(nn.Linear(d, h))
            d = h
        self.out_dim = d


class GAMMAPrimitiveVAE(nn.Module):
    def __init__(self, configs):
        super().__init__()
        if configs.get("body_repr", "ssm2_67") != "ssm2_67":
            raise NotImplementedError("crowd_ppo uses body_repr ssm2_67 for the predictor")
        self.in_dim = in_dim = 67 * 3
        self.h_dim = h = configs["h_dim"]
        self.z_dim = z = configs["z_dim"]
        hd = list(configs["hdims_mlp"])
        if not configs.get("use_drnn_mlp", True) or not configs.get("residual", True) or len(hd) != 2:
            raise NotImplementedError("only the use_drnn_mlp + residual + 2-layer MLP configuration is built")
        self.x_enc = nn.GRU(in_dim, h)
        self.e_rnn = nn.GRU(in_dim, h)
        self.e_mlp = _MLP(2 * h, hd)
        self.e_mu = nn.Linear(hd[-1], z)
        self.e_logvar = nn.Linear(hd[-1], z)
        self.drnn_mlp = _MLP(h, hd + [h])
        self.d_rnn = nn.GRUCell(in_dim + z + h, h)
        self.d_mlp = _MLP(h, hd)
        self.d_out = nn.Linear(hd[-1], in_dim)


class _ResNetBlock(nn.Module):
    def __init__(self, in_dim, h_dim, out_dim, n_blocks):
        super().__init__()
        self.in_fc = nn.Linear(in_dim, h_dim)
        self.layers = nn.ModuleList([_MLP(h_dim, (h_dim, h_dim)) for _ in range(n_blocks)])
        self.out_fc = nn.Linear(h_dim, out_dim)


class MoshRegressor(nn.Module):
    def __init__(self, config):
        super().__init__()
        if not config.get("use_cont", False) or config.get("actfun", "relu") != "relu":
            raise NotImplementedError("crowd_ppo uses the use_cont + relu regressor (MoshRegressor_v3_male.yml)")
        self.in_dim = 67 * 3
        self.h_dim, self.n_blocks, self.n_recur = config["h_dim"], config["n_blocks"], config["n_recur"]
        self.body_dim = 3 + 6 + 21 * 6 + 24
        self.pnet = _ResNetBlock(self.in_dim + self.body_dim + 10, self.h_dim, self.body_dim, self.n_blocks)


PREDICTOR_CFG = {"body_repr": "ssm2_67", "h_dim": 256, "z_dim": 128, "t_his": 2, "t_pred": 18,
                 "use_drnn_mlp": True, "hdims_mlp": [512, 256], "residual": True}   # MPVAE_samp20_2frame_rollout.yml
REGRESSOR_CFG = {"gender": "male", "h_dim": 128, "n_blocks": 10, "n_recur": 3, "body_repr": "ssm2_67",
                 "actfun": "relu", "use_cont": True}                                # MoshRegressor_v3_male.yml


class GAMMAPrimitiveCombo(nn.Module):
    """predictor + regressor; ``sample_prior`` keeps the reference signature (:334-360)."""

    def __init__(self, markercfg=None, bparamscfg=None):
        super().__init__()
        self.predictor = GAMMAPrimitiveVAE(markercfg or PREDICTOR_CFG)
        self.regressor = MoshRegressor(bparamscfg or REGRESSOR_CFG)
        self._h = None
        self._wkeep = None

    # ---- C handle ------------------------------------------------------------------------
    def _weight_list(self):
        p, r = self.predictor, self.regressor
        ws = [p.x_enc.weight_ih_l0, p.x_enc.weight_hh_l0, p.x_enc.bias_ih_l0, p.x_enc.bias_hh_l0]
        for l in p.drnn_mlp.layers:
            ws += [l.weight, l.bias]
        ws += [p.d_rnn.weight_ih, p.d_rnn.weight_hh, p.d_rnn.bias_ih, p.d_rnn.bias_hh]
        for l in p.d_mlp.layers:
            ws += [l.weight, l.bias]
        ws += [p.d_out.weight, p.d_out.bias, r.pnet.in_fc.weight, r.pnet.in_fc.bias]
        for blk in r.pnet.layers:
            for l in blk.layers:
                ws += [l.weight, l.bias]
        ws += [r.pnet.out_fc.weight, r.pnet.out_fc.bias]
        return ws

    def handle(self):
        ws = self._weight_list()
        key = tuple(w.data_ptr() for w in ws)
        if self._h is None or key != self._wkeep:
            self.release()
            dev = ws[0].device
            if dev.type != "cuda":
                raise _lib.EgError("motion model parameters must live on a CUDA device (no CPU path)")
            for w in ws:
                if not w.is_contiguous() or w.dtype != torch.float32 or w.device != dev:
                    raise _lib.EgError("motion model parameters must be contiguous float32 on one device")
            p, r = self.predictor, self.regressor
            dims = _lib.EgMotionDims(p.in_dim, p.h_dim, p.z_dim, p.d_mlp.layers[0].out_features, r.h_dim,
                                     r.n_blocks, r.n_recur, r.body_dim)
            arr = (C.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
            h = C.c_void_p()
            _lib.check(_lib.lib().eg_motion_create(C.byref(dims), arr, len(ws), dev.index or 0, C.byref(h)))
            self._h, self._wkeep = h, key
        return self._h

    def refresh_weights(self):
        """Re-derive the library's transposed weight copies after an in-place weight change."""
        if self._h is not None:
            _lib.check(_lib.lib().eg_motion_refresh(self._h, None))
            torch.cuda.synchronize()

    def set_fused(self, fused: bool):
        _lib.check(_lib.lib().eg_motion_set_fused(self.handle(), int(bool(fused))))

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.refresh_weights()
        return r

    def release(self):
        if getattr(self, "_h", None) is not None:
            _lib.lib().eg_motion_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # ---- operators -----------------------------------------------------------------------
    def sample_prior_env_major(self, X, ldx_env, ldx_frame, z, betas, B):
        """X: CUDA tensor whose frame t of env b starts at element b*ldx_env + t*ldx_frame.
        Returns Y [B,20,201] (history + prediction) and Yb [B,20,93] (frames 2.. valid)."""
        dev = z.device
        Y = torch.empty(B, 20, 201, dtype=torch.float32, device=dev)
        Yb = torch.zeros(B, 20, 93, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().eg_motion_sample_prior(self.handle(), _lib.ptr(X), ldx_env, ldx_frame, _lib.ptr(z),
                                                         _lib.ptr(betas), B, _lib.ptr(Y), _lib.ptr(Yb),
                                                         _lib.stream_ptr(dev)))
        return Y, Yb

    def sample_prior(self, X, betas, z=None):
        """X [t_his=2, b, 201], betas [18, b, 10] (or [b,10]), z [b,128] -> (Y_gen [18,b,201], Yb_gen [18,b,93])."""
        if X.shape[0] != 2:
            raise NotImplementedError("the 2-frame motion seed model is the one crowd_ppo uses")
        b = X.shape[1]
        if z is None:
            z = torch.randn((b, self.predictor.z_dim), device=X.device)
        Xe = X.to(torch.float32).permute(1, 0, 2).contiguous()            # [b,2,201]
        be = betas.reshape(-1, b, 10)[0] if betas.dim() == 3 else betas.reshape(b, 10)
        Y, Yb = self.sample_prior_env_major(Xe, 2 * 201, 201, z.to(torch.float32).contiguous(),
                                            be.to(torch.float32).contiguous(), b)
        return Y[:, 2:].permute(1, 0, 2).contiguous(), Yb[:, 2:].permute(1, 0, 2).contiguous()


class GAMMAPrimitiveComboGenOP:
    """Loader with the reference's interface (:1099-1139): ``.model`` is the combo; ``build_model`` reads
    ``<predictor save_dir>/epoch-400.ckp`` (fallback 200) and ``<regressor save_dir>/epoch-100.ckp`` when
    present, else fills seeded synthetic weights (no checkpoints exist offline)."""

    def __init__(self, predictorcfg=None, regressorcfg=None, testconfig=None):
        self.predictorcfg, self.regressorcfg, self.testconfig = predictorcfg, regressorcfg, testconfig or {}
        self.device = torch.device("cuda", self.testconfig.get("gpu_index", 0))
        self.t_his = 2
        self.model = None

    def build_model(self, load_pretrained_model=False, predictor_dir=None, regressor_dir=None, seed=0):
        """load_pretrained_model=True reads ``<predictor_dir>/epoch-400.ckp`` (else epoch-200.ckp) and
        ``<regressor_dir>/epoch-100.ckp`` like the reference (:1118-1134) and RAISES when one is missing; False fills
        seeded synthetic weights and says so loudly (offline tests / benchmarks: no licensed checkpoints exist here)."""
        self.model = GAMMAPrimitiveCombo()
        if load_pretrained_model:
            if not (predictor_dir and regressor_dir):
                raise FileNotFoundError("load_pretrained_model=True needs predictor_dir and regressor_dir")
            cand = [os.path.join(predictor_dir, n) for n in ("epoch-400.ckp", "epoch-200.ckp")]
            ppath = next((c for c in cand if os.path.exists(c)), None)
            rpath = os.path.join(regressor_dir, "epoch-100.ckp")
            if ppath is None or not os.path.exists(rpath):
                raise FileNotFoundError(f"motion-model checkpoints not found: {cand[0]} (or epoch-200.ckp) and {rpath}")
            self.model.predictor.load_state_dict(torch.load(ppath, map_location="cpu")["model_state_dict"])
            self.model.regressor.load_state_dict(torch.load(rpath, map_location="cpu")["model_state_dict"])
            self.weights = {"predictor": ppath, "regressor": rpath}
        else:
            import warnings
            warnings.warn("GAMMAPrimitiveComboGenOP: using seeded SYNTHETIC motion-model weights (no pretrained GAMMA "
                          "checkpoints were requested); pass predictor_dir / regressor_dir for the real model", stacklevel=2)
            self.weights = "synthetic"
            assets.fill_params_(self.model.predictor, seed=seed + 11)
            assets.fill_params_(self.model.regressor, seed=seed + 12, w_gain=0.7)
            with torch.no_grad():   # keep the synthetic primitives human-scale (cm per frame, small rotations)
                self.model.predictor.d_out.weight.mul_(0.02); self.model.predictor.d_out.bias.mul_(0.02)
                self.model.regressor.pnet.out_fc.weight.mul_(0.3)
        self.model.eval()
        self.model.to(self.device)
        return self.model


class VPoserEncoder(nn.Module):
    """human_body_prior VPoser v1.0 encoder parameters (state_dict names as in vposer_v1_0 snapshots);
    ``encode(x).loc`` mirrors the reference call site crowd_env_2f.py:198."""

    def __init__(self, num_neurons=512, latentD=32, n_features=63):
        super().__init__()
        self.bodyprior_enc_bn1 = nn.BatchNorm1d(n_features)
        self.bodyprior_enc_fc1 = nn.Linear(n_features, num_neurons)
        self.bodyprior_enc_bn2 = nn.BatchNorm1d(num_neurons)
        self.bodyprior_enc_fc2 = nn.Linear(num_neurons, num_neurons)
        self.bodyprior_enc_mu = nn.Linear(num_neurons, latentD)
        self.bodyprior_enc_logvar = nn.Linear(num_neurons, latentD)
        self._h = None
        self._wkeep = None
        self.eval()

    def _weight_list(self):
        b1, b2 = self.bodyprior_enc_bn1, self.bodyprior_enc_bn2
        return [b1.weight, b1.bias, b1.running_mean, b1.running_var, self.bodyprior_enc_fc1.weight,
                self.bodyprior_enc_fc1.bias, b2.weight, b2.bias, b2.running_mean, b2.running_var,
                self.bodyprior_enc_fc2.weight, self.bodyprior_enc_fc2.bias, self.bodyprior_enc_mu.weight,
                self.bodyprior_enc_mu.bias]

    def handle(self):
        ws = self._weight_list()
        key = tuple(w.data_ptr() for w in ws)
        if self._h is None or key != self._wkeep:
            if self._h is not None:
                _lib.lib().eg_vposer_destroy(self._h)
            dev = ws[0].device
            if dev.type != "cuda":
                raise _lib.EgError("VPoser parameters must live on a CUDA device (no CPU path)")
            arr = (C.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
            h = C.c_void_p()
            _lib.check(_lib.lib().eg_vposer_create(arr, len(ws), dev.index or 0, C.byref(h)))
            self._h, self._wkeep = h, key
        return self._h

    def encode(self, pin):
        from types import SimpleNamespace
        x = pin.reshape(pin.shape[0], -1).to(torch.float32).contiguous()
        loc = torch.empty(x.shape[0], 32, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().eg_vposer_encode(self.handle(), _lib.ptr(x), x.shape[1], x.shape[0], _lib.ptr(loc),
                                                   _lib.stream_ptr(x.device)))
        return SimpleNamespace(loc=loc)


def load_vposer(expr_dir=None, vp_model="snapshot", seed=0, device="cuda"):
    """Signature of human_body_prior.tools.model_loader.load_vposer (main_ppo.py:259): returns (vposer, cfg).
    Reads ``<expr_dir>/snapshots/*.pt`` when present, else seeded synthetic weights."""
    import glob
    vp = VPoserEncoder()
    snaps = sorted(glob.glob(os.path.join(expr_dir, "snapshots", "*.pt"))) if expr_dir else []
    if snaps:
        sd = torch.load(snaps[-1], map_location="cpu")
        missing = vp.load_state_dict({k: v for k, v in sd.items() if k.startswith("bodyprior_enc")}, strict=False)
        need = [k for k in missing.missing_keys if "logvar" not in k and "num_batches_tracked" not in k]
        if need:
            raise KeyError(f"VPoser snapshot {snaps[-1]} lacks encoder tensors: {need}")
    elif expr_dir:
        raise FileNotFoundError(f"no VPoser snapshot under {os.path.join(expr_dir, 'snapshots')}")
    else:
        import warnings
        warnings.warn("load_vposer: using seeded SYNTHETIC VPoser encoder weights (no expr_dir given)", stacklevel=2)
        assets.fill_params_(vp, seed=seed + 31)
    vp.eval()
    return vp.to(device), None
