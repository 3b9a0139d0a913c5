"""GAMMAPPOPolicy - host-side mirror of the reference's motion/crowd_ppo/ppo_policy.py (which extends
tianshou 0.5.0's PPOPolicy). Same constructor keywords, same ``forward`` / ``learn`` surface and the
same checkpoint layout (state_dict groups actor.*, critic.*, _actor_critic.actor.*,
_actor_critic.critic.*, shared_net.*); the arithmetic (network forward, GAE, PPO loss, backward,
grad-norm clip, AdamW) is in egogen_b200/csrc/ppo.cu. torch is used for buffers, RNG, the NCCL
allreduce (torch.distributed) and nothing else.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from .models_policy_ppo import ActorCritic

_EPS = float(np.finfo(np.float32).eps)     # tianshou BasePolicy._eps


class Batch(SimpleNamespace):
    """Minimal stand-in for tianshou.data.Batch (attribute bag)."""


class GAMMAPPOPolicy(nn.Module):
    def __init__(self, actor, critic, shared_net, optim, dist_fn=None, eps_clip=0.2, weight_kld=1.0, dual_clip=None,
                 value_clip=False, advantage_normalization=True, recompute_advantage=False, discount_factor=0.99,
                 gae_lambda=0.95, max_grad_norm=None, vf_coef=0.5, ent_coef=0.01, reward_normalization=False,
                 max_batchsize=256, deterministic_eval=False, action_space=None, action_scaling=False,
                 action_bound_method="", lr_scheduler=None, process_group=None, **kwargs):
        super().__init__()
        if dual_clip is not None or value_clip or recompute_advantage or reward_normalization:
            raise NotImplementedError("dual_clip / value_clip / recompute_adv / rew_norm are off in main_ppo.py:54-67")
        if action_scaling or action_bound_method:
            raise NotImplementedError("main_ppo.py:152-155 disables action scaling / bounding")
        # registration order reproduces the reference's state_dict key order
        self.actor, self.critic = actor, critic
        self._actor_critic = ActorCritic(actor, critic)          # tianshou's 2-module ActorCritic (quirk 1)
        self.shared_net = shared_net
        self.optim, self.dist_fn = optim, dist_fn
        self._eps_clip, self._weight_kld, self._norm_adv = eps_clip, weight_kld, bool(advantage_normalization)
        self._gamma, self._lambda = discount_factor, gae_lambda
        self._grad_norm, self._weight_vf, self._weight_ent = max_grad_norm, vf_coef, ent_coef
        self._batch, self._deterministic_eval = max_batchsize, deterministic_eval
        self._eps = _EPS
        self.pg = process_group
        self._flatten()
        self._h = None
        self._opt_step = 0
        self.gen = None                                            # optional torch.Generator for action noise

    # ---- flat parameter / gradient / moment buffers ------------------------------------------
    def _ordered_params(self):
        return list(self.actor.parameters()) + list(self.critic.parameters()) + list(self.shared_net.parameters())

    def _flatten(self):
        ps = self._ordered_params()
        dev = ps[0].device
        if dev.type != "cuda":
            raise _lib.EgError("policy parameters must live on a CUDA device (no CPU path)")
        n = sum(p.numel() for p in ps)
        self.dims = _lib.EgPolicyDims(self.shared_net.in_dim, 32, self.shared_net.h_dim, 32, self.actor.n_blocks,
                                      self.actor.z_dim)
        nac = C.c_int64()
        expect = _lib.lib().eg_policy_param_count(C.byref(self.dims), C.byref(nac))
        if expect != n:
            raise _lib.EgError(f"parameter count {n} does not match the library layout {expect}")
        self.n_actor_critic = nac.value
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
        self._stats = torch.zeros(8, dtype=torch.float32, device=dev)
        self._mom = torch.zeros(3, dtype=torch.float64, device=dev)
        self.dev = dev

    def handle(self):
        if self._h is None:
            h = C.c_void_p()
            _lib.check(_lib.lib().eg_policy_create(C.byref(self.dims), _lib.ptr(self.flat_params),
                                                   _lib.ptr(self.flat_grads), self.dev.index or 0, C.byref(h)))
            self._h = h
        return self._h

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.lib().eg_policy_destroy(self._h)
        except Exception:
            pass

    # ---- network forward ---------------------------------------------------------------------
    def _obs_ptrs(self, obs):
        st = _lib.f32c(obs["state"], self.dev)
        eg = _lib.f32c(obs["egosensing"], self.dev)
        di = _lib.f32c(obs["dist"], self.dev).reshape(-1)
        ti = _lib.f32c(obs["time"], self.dev).reshape(-1)
        return st, eg, di, ti

    def net_forward(self, obs, want_actor=True, want_critic=True):
        """-> (out_actor [B,256] raw [mu|logvar] or None, value [B] or None)"""
        st, eg, di, ti = self._obs_ptrs(obs)
        B = st.shape[0]
        oa = torch.empty(B, 2 * self.actor.z_dim, dtype=torch.float32, device=self.dev) if want_actor else None
        val = torch.empty(B, dtype=torch.float32, device=self.dev) if want_critic else None
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_policy_forward(self.handle(), _lib.ptr(st), _lib.ptr(eg), _lib.ptr(di), _lib.ptr(ti),
                                                    B, int(want_actor), int(want_critic), _lib.ptr(oa), _lib.ptr(val),
                                                    None, _lib.stream_ptr(self.dev)))
        return oa, val

    def forward(self, batch, state=None, want_value=False, **kwargs):
        """ppo_policy.py:142-179. ``batch.obs`` is the observation dict. Returns Batch(logits=(mu, sigma), act,
        state=None, z_mu, z_var, z_logvar, logp[, value])."""
        obs = batch.obs if hasattr(batch, "obs") else batch["obs"]
        oa, val = self.net_forward(obs, True, want_value)
        Z = self.actor.z_dim
        B = oa.shape[0]
        deterministic = self._deterministic_eval and not self.training
        eps = None if deterministic else torch.randn(B, Z, device=self.dev, generator=self.gen)
        act = torch.empty(B, Z, dtype=torch.float32, device=self.dev)
        logp = torch.empty(B, dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_gauss_sample(_lib.ptr(oa), _lib.ptr(eps), B, Z, float(self.actor.min_logvar),
                                                  float(self.actor.max_logvar), _lib.ptr(act), _lib.ptr(logp),
                                                  _lib.stream_ptr(self.dev)))
        z_mu = oa[:, :Z]
        z_logvar = oa[:, Z:].clamp(self.actor.min_logvar, self.actor.max_logvar)
        z_var = torch.exp(z_logvar)
        return Batch(logits=(z_mu, z_var ** 0.5), act=act, state=None, z_mu=z_mu, z_var=z_var, z_logvar=z_logvar,
                     logp=logp, value=val)

    # ---- returns / advantages ----------------------------------------------------------------
    def compute_returns(self, v_s, v_next, rew, terminated, end_flag):
        """_compute_returns (:105-140) + tianshou compute_episodic_return on [T,E] time-major tensors."""
        T, E = rew.shape
        adv = torch.empty(T, E, dtype=torch.float32, device=self.dev)
        ret = torch.empty(T, E, dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_gae(_lib.ptr(v_s.contiguous()), _lib.ptr(v_next.contiguous()),
                                         _lib.ptr(rew.contiguous()), _lib.ptr(terminated.contiguous()),
                                         _lib.ptr(end_flag.contiguous()), T, E, float(self._gamma), float(self._lambda),
                                         _lib.ptr(adv), _lib.ptr(ret), _lib.stream_ptr(self.dev)))
        return ret, adv

    # ---- learning ----------------------------------------------------------------------------
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.pg) if (dist.is_available() and dist.is_initialized()) else 1

    def normalize_adv(self, adv):
        """per-minibatch (mean, unbiased std) normalisation (:192-195); global over ranks when distributed."""
        lib = _lib.lib()
        n = adv.numel()
        with torch.cuda.device(self.dev):
            _lib.check(lib.eg_moments(_lib.ptr(adv), n, _lib.ptr(self._mom), _lib.stream_ptr(self.dev)))
            self._mom[2] = float(n)
            if self._world() > 1:
                import torch.distributed as dist
                dist.all_reduce(self._mom, group=self.pg)
            out = torch.empty_like(adv)
            _lib.check(lib.eg_adv_normalize(_lib.ptr(adv), n, _lib.ptr(self._mom), self._eps, _lib.ptr(out),
                                            _lib.stream_ptr(self.dev)))
        return out

    def _opt_hparams(self):
        g = self.optim.param_groups[0]
        b1, b2 = g.get("betas", (0.9, 0.999))
        return float(g["lr"]), float(b1), float(b2), float(g.get("eps", 1e-8)), float(g.get("weight_decay", 0.01))

    def loss_backward(self, mb, global_batch: Optional[int] = None):
        """forward + PPO loss + backward of one minibatch into flat_grads; stats land in self._stats."""
        B = mb.act.shape[0]
        gb = global_batch if global_batch is not None else B * self._world()
        adv = self.normalize_adv(mb.adv.contiguous()) if self._norm_adv else mb.adv.contiguous()
        st, eg, di, ti = self._obs_ptrs(mb.obs)
        self._stats.zero_()
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_ppo_loss_backward(
                self.handle(), _lib.ptr(st), _lib.ptr(eg), _lib.ptr(di), _lib.ptr(ti), _lib.ptr(mb.act.contiguous()),
                _lib.ptr(mb.logp_old.contiguous()), _lib.ptr(adv), _lib.ptr(mb.returns.contiguous()), B, 1.0 / gb,
                float(self._eps_clip), float(self._weight_vf), float(self._weight_ent), float(self.actor.min_logvar),
                float(self.actor.max_logvar), 1, _lib.ptr(self._stats), _lib.stream_ptr(self.dev)))

    def optimizer_step(self):
        lr, b1, b2, eps, wd = self._opt_hparams()
        self._opt_step += 1
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_clip_adamw_step(self.handle(), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                                     float(self._grad_norm or 0.0), lr, b1, b2, eps, wd, self._opt_step,
                                                     _lib.stream_ptr(self.dev)))

    def learn_minibatch(self, mb, global_batch: Optional[int] = None):
        """One iteration of the inner loop of learn (:189-252). mb: Batch(obs, act, logp_old, adv, returns)
        with CUDA tensors. Returns a clone of the device stats tensor."""
        self.loss_backward(mb, global_batch)
        if self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_grads, group=self.pg)          # ONE allreduce of the whole gradient
            dist.all_reduce(self._stats, group=self.pg)
        self.optimizer_step()
        return self._stats.clone()

    def learn(self, batch, batch_size: int, repeat: int, **kwargs):
        """learn (:182-265). batch: Batch of flat [N,...] CUDA tensors (obs dict, act, logp_old, adv, returns).
        batch_size is the PER-RANK minibatch size (global = batch_size * world). Minibatch order comes from
        np.random.permutation like tianshou's Batch.split(shuffle=True, merge_last=True)."""
        N = batch.act.shape[0]
        out = {k: [] for k in ("loss", "loss/clip", "loss/vf", "loss/ent", "loss/kld")}
        all_stats = []
        for step in range(repeat):
            perm = np.random.permutation(N)
            starts = list(range(0, N, batch_size))
            if len(starts) > 1 and N - starts[-1] < batch_size:
                starts = starts[:-1]                                   # merge_last
            stats = None
            for i, s in enumerate(starts):
                e = N if i == len(starts) - 1 else s + batch_size
                idx = torch.as_tensor(perm[s:e], device=self.dev)
                mb = Batch(obs={k: v.index_select(0, idx) for k, v in batch.obs.items()},
                           act=batch.act.index_select(0, idx), logp_old=batch.logp_old.index_select(0, idx),
                           adv=batch.adv.index_select(0, idx), returns=batch.returns.index_select(0, idx))
                stats = self.learn_minibatch(mb)
                all_stats.append(stats)
            if repeat > 1 and stats is not None and float(stats[4].item()) >= 0.02:
                break                                                  # KL early stop on the last minibatch (:254-257)
        if all_stats:
            S = torch.stack(all_stats).cpu().numpy()                   # one D2H read per learn() call
            for s in S:
                clip, vf, ent, kld = float(s[0]), float(s[1]), float(s[2]), float(s[3])
                out["loss/clip"].append(clip); out["loss/vf"].append(vf); out["loss/ent"].append(ent)
                out["loss/kld"].append(kld)
                out["loss"].append(clip + self._weight_vf * vf - self._weight_ent * ent)
        return out

    # ---- optimiser state in torch.optim.AdamW format (checkpoint "optim" entry, main_ppo.py:207-213) ----
    def export_optim_state(self):
        off = 0
        for p in self._ordered_params():
            k = p.numel()
            self.optim.state[p] = {"step": torch.tensor(float(self._opt_step)),
                                   "exp_avg": self.exp_avg[off:off + k].view_as(p),
                                   "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p)}
            off += k
        return self.optim.state_dict()
