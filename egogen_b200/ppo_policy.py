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


class PolicyOutput(Batch):
    """Return value of GAMMAPPOPolicy.forward. ``dist`` is the action distribution of the reference
    (``dist_fn(*logits)``, main_ppo.py:148-149: Independent(Normal(mu, sigma), 1)); it is built on first access so the
    collect loop, which only needs ``act`` / ``logp``, does not pay for a torch distribution object per step."""

    def __init__(self, dist_fn=None, **kw):
        super().__init__(**kw)
        self._dist_fn, self._dist = dist_fn, None

    @property
    def dist(self):
        if self._dist is None:
            if self._dist_fn is not None:
                self._dist = self._dist_fn(*self.logits)
            else:
                from torch.distributions import Independent, Normal
                self._dist = Independent(Normal(*self.logits), 1)
        return self._dist


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
        self.dims = _lib.EgPolicyDims(self.shared_net.in_dim, 32, self.shared_net.h_dim, 32, self.actor.n_blocks,
                                      self.actor.z_dim)
        nac = C.c_int64()
        n = _lib.lib().eg_policy_param_count(C.byref(self.dims), C.byref(nac))       # flat length incl. alignment padding
        offs = (C.c_int64 * 64)()
        nt = _lib.lib().eg_policy_param_offsets(C.byref(self.dims), offs, 64)
        if nt != len(ps):
            raise _lib.EgError(f"{len(ps)} parameter tensors do not match the library layout ({nt})")
        self._offsets = [int(offs[i]) for i in range(nt)]
        for i, p in enumerate(ps):                                               # the layout must hold every tensor
            end = self._offsets[i + 1] if i + 1 < nt else n
            if self._offsets[i] + p.numel() > end:
                raise _lib.EgError(f"parameter tensor {i} ({tuple(p.shape)}) does not fit the library layout")
        self.n_actor_critic = nac.value
        self.n_params = n
        self._dp = None
        if self._world() > 1:
            self._dp = self._setup_nvlink_optimizer(n, dev)
        if self._dp is not None:
            self.flat_params, self.flat_grads = self._dp["params"][:n], self._dp["grads"][:n]
            chunk = self._dp["n_pad"] // self._dp["world"]
            self.exp_avg = torch.zeros(chunk, dtype=torch.float32, device=dev)        # this rank's moment slices
            self.exp_avg_sq = torch.zeros(chunk, dtype=torch.float32, device=dev)
        else:
            self.flat_params = torch.zeros(n, dtype=torch.float32, device=dev)
            self.flat_grads = torch.zeros(n, dtype=torch.float32, device=dev)
            self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
            self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        for off, p in zip(self._offsets, ps):
            k = p.numel()
            self.flat_params[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + k].view_as(p)
            p.grad = self.flat_grads[off:off + k].view_as(p)
        self._stats = torch.zeros(8, dtype=torch.float32, device=dev)
        self._mom = torch.zeros(3, dtype=torch.float64, device=dev)
        self.dev = dev
        if self._world() > 1:                                      # every replica starts from rank 0's initialisation
            import torch.distributed as dist
            dist.broadcast(self.flat_params, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)

    def _setup_nvlink_optimizer(self, n, dev):
        """Flat parameter / gradient vectors in torch symmetric memory (peer-mapped over NVLink, multicast-mapped through
        the NVSwitch when the fabric offers it) for the sharded optimiser step of csrc/dp_optim.cu. Returns None - and the
        policy falls back to one NCCL all_reduce of the gradient per step - when symmetric memory cannot be set up
        (EG_DP_OPTIM=0, a non-NCCL group such as the gloo CPU tests, or no peer access)."""
        import os
        import torch.distributed as dist
        if os.environ.get("EG_DP_OPTIM", "1") == "0":
            return None
        try:
            group = self.pg if self.pg is not None else dist.group.WORLD
            if dist.get_backend(group) != "nccl":
                return None
            import torch.distributed._symmetric_memory as symm_mem
            try:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    symm_mem.enable_symm_mem_for_group(group.group_name)      # required by older torch, a no-op since
            except Exception:                                     # noqa: BLE001
                pass
            W, rank = dist.get_world_size(group), dist.get_rank(group)
            if W > 16:
                return None
            n_pad = -(-n // (4 * W)) * 4 * W
            bufs = {"params": symm_mem.empty(n_pad, dtype=torch.float32, device=dev),
                    "grads": symm_mem.empty(n_pad, dtype=torch.float32, device=dev),
                    "scratch": symm_mem.empty(max(W, 2), dtype=torch.float64, device=dev)}
            hdl = {}
            for k, t in bufs.items():
                t.zero_()
                hdl[k] = symm_mem.rendezvous(t, group.group_name)
            use_mc = os.environ.get("EG_DP_MULTICAST", "1") != "0"
            mc = {k: (int(getattr(hdl[k], "multicast_ptr", 0) or 0) if use_mc else 0) for k in ("params", "grads")}
            ptrs = {k: (C.c_void_p * W)(*[int(x) for x in hdl[k].buffer_ptrs]) for k in bufs}
            dp = dict(bufs, hdl=hdl, mc=mc, ptrs=ptrs, world=W, rank=rank, n_pad=n_pad,
                      gred=torch.zeros(n_pad // W, dtype=torch.float32, device=dev),
                      work=torch.zeros(2, dtype=torch.float64, device=dev))
            torch.cuda.synchronize(dev)
            hdl["grads"].barrier(channel=0)
            return dp
        except Exception as e:                                     # noqa: BLE001 - any failure means "no symmetric memory here"
            import warnings
            warnings.warn(f"NVLink sharded optimiser unavailable ({type(e).__name__}: {e}); using NCCL all_reduce")
            return None

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
        return PolicyOutput(self.dist_fn, logits=(z_mu, z_var ** 0.5), act=act, state=None, z_mu=z_mu, z_var=z_var,
                            z_logvar=z_logvar, logp=logp, value=val)

    # ---- returns / advantages ----------------------------------------------------------------
    def process_fn(self, batch, buffer=None, indices=None):
        """ppo_policy.py:93-140 (process_fn + _compute_returns): critic values of obs / obs_next, GAE returns and
        advantages, old log-probabilities. ``batch``: obs, obs_next (observation dicts), act, rew, terminated, truncated as
        flat [N] tensors in tianshou's buffer order (env-major: env 0's T steps, then env 1's, ...); ``buffer`` only has to
        say how many envs that is (``buffer.E`` / ``buffer.buffer_num``; default: one trajectory). The last stored step of
        every env closes its segment like tianshou's unfinished_index(). Adds v_s, returns, adv, logp_old to ``batch``."""
        N = batch.act.shape[0]
        E = int(getattr(buffer, "E", getattr(buffer, "buffer_num", 1)) or 1)
        T = N // E
        to = lambda x: torch.as_tensor(x, device=self.dev)
        with torch.no_grad():
            chunks = [(s, min(N, s + self._batch)) for s in range(0, N, self._batch)]
            if len(chunks) > 1 and chunks[-1][1] - chunks[-1][0] < self._batch:      # split(..., merge_last=True)
                chunks[-2:] = [(chunks[-2][0], N)]
            v_s = torch.cat([self.net_forward({k: to(v)[a:b] for k, v in batch.obs.items()}, False, True)[1] for a, b in chunks])
            v_n = torch.cat([self.net_forward({k: to(v)[a:b] for k, v in batch.obs_next.items()}, False, True)[1] for a, b in chunks])
            tm = lambda x: to(x).reshape(E, T).t().contiguous()                          # env-major flat -> [T,E]
            term = tm(batch.terminated).to(torch.uint8)
            end = (term.bool() | tm(batch.truncated).bool()).to(torch.uint8)
            end[-1] = 1
            ret, adv = self.compute_returns(tm(v_s), tm(v_n), tm(batch.rew).to(torch.float32), term, end)
            fl = lambda x: x.t().reshape(N).contiguous()
            batch.v_s, batch.returns, batch.adv = v_s, fl(ret), fl(adv)
            batch.act = to(batch.act).to(torch.float32)
            batch.logp_old = self(Batch(obs={k: to(v) for k, v in batch.obs.items()})).dist.log_prob(batch.act)
        return batch

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

    def adv_moments(self, adv, out=None):
        """{sum, sum of squares, count} of one minibatch's advantages as device doubles (local to this rank)."""
        mom = self._mom if out is None else out
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_moments(_lib.ptr(adv), adv.numel(), _lib.ptr(mom), _lib.stream_ptr(self.dev)))
        mom[2] = float(adv.numel())
        return mom

    def normalize_adv(self, adv, moments=None):
        """per-minibatch (mean, unbiased std) normalisation (:192-195); global over ranks when distributed.
        `moments`: already rank-summed {sum, sumsq, count} (learn() reduces the moments of ALL its minibatches in one
        collective up front); None computes and reduces them here."""
        n = adv.numel()
        if moments is None:
            moments = self.adv_moments(adv)
            if self._world() > 1:
                import torch.distributed as dist
                dist.all_reduce(moments, group=self.pg)
        out = torch.empty_like(adv)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_adv_normalize(_lib.ptr(adv), n, _lib.ptr(moments), self._eps, _lib.ptr(out),
                                                   _lib.stream_ptr(self.dev)))
        return out

    def _opt_hparams(self):
        g = self.optim.param_groups[0]
        b1, b2 = g.get("betas", (0.9, 0.999))
        return float(g["lr"]), float(b1), float(b2), float(g.get("eps", 1e-8)), float(g.get("weight_decay", 0.01))

    def loss_backward(self, mb, global_batch: Optional[int] = None, moments=None, part: str = "all"):
        """forward + PPO loss + backward of one minibatch into flat_grads; stats land in self._stats.
        part = "mlp" stops after the actor / critic chains (eg_ppo_loss_backward_mlp: every gradient of the actor + critic
        prefix is final), part = "encoders" finishes the same minibatch (eg_ppo_backward_encoders)."""
        if part == "encoders":
            eg, B = self._pending
            with torch.cuda.device(self.dev):
                _lib.check(_lib.lib().eg_ppo_backward_encoders(self.handle(), _lib.ptr(eg), B, _lib.stream_ptr(self.dev)))
            self._pending = None
            return
        B = mb.act.shape[0]
        gb = global_batch if global_batch is not None else B * self._world()
        adv = self.normalize_adv(mb.adv.contiguous(), moments) if self._norm_adv else mb.adv.contiguous()
        st, eg, di, ti = self._obs_ptrs(mb.obs)
        self._stats.zero_()
        fn = _lib.lib().eg_ppo_loss_backward_mlp if part == "mlp" else _lib.lib().eg_ppo_loss_backward
        with torch.cuda.device(self.dev):
            _lib.check(fn(
                self.handle(), _lib.ptr(st), _lib.ptr(eg), _lib.ptr(di), _lib.ptr(ti), _lib.ptr(mb.act.contiguous()),
                _lib.ptr(mb.logp_old.contiguous()), _lib.ptr(adv), _lib.ptr(mb.returns.contiguous()), B, 1.0 / gb,
                float(self._eps_clip), float(self._weight_vf), float(self._weight_ent), float(self.actor.min_logvar),
                float(self.actor.max_logvar), 1, _lib.ptr(self._stats), _lib.stream_ptr(self.dev)))
        self._pending = (eg, B) if part == "mlp" else None

    def optimizer_step(self):
        """clip_grad_norm_(actor + critic) + AdamW (:241-247). With N > 1 ranks the gradient is first summed over the
        ranks: by the sharded NVLink step (eg_dp_reduce_norm -> eg_dp_adamw_gather, each rank updates 1/N of the
        parameters and broadcasts them) when symmetric memory is set up, else by one NCCL all_reduce."""
        lr, b1, b2, eps, wd = self._opt_hparams()
        self._opt_step += 1
        lib, st = _lib.lib(), _lib.stream_ptr(self.dev)
        dp = self._dp
        with torch.cuda.device(self.dev):
            if dp is not None:
                W, rank, n_pad = dp["world"], dp["rank"], dp["n_pad"]
                bar = dp["hdl"]["grads"].barrier
                bar(channel=0)                                       # every rank's backward has written its gradient
                _lib.check(lib.eg_dp_reduce_norm(dp["ptrs"]["grads"], C.c_void_p(dp["mc"]["grads"] or None), W, rank, n_pad,
                                                 self.n_actor_critic, _lib.ptr(dp["gred"]), dp["ptrs"]["scratch"],
                                                 _lib.ptr(dp["work"]), st))
                bar(channel=0)                                       # all norm shares published, all slices read
                _lib.check(lib.eg_dp_adamw_gather(dp["ptrs"]["params"], C.c_void_p(dp["mc"]["params"] or None), W, rank, n_pad,
                                                  self.n_actor_critic, _lib.ptr(dp["gred"]), _lib.ptr(dp["scratch"]),
                                                  _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                                  float(self._grad_norm or 0.0), lr, b1, b2, eps, wd, self._opt_step, st))
                bar(channel=0)                                       # every rank's slice of the new parameters has landed
                return
            if self._world() > 1:
                import torch.distributed as dist
                dist.all_reduce(self.flat_grads, group=self.pg)      # fallback: ONE allreduce of the whole gradient
            _lib.check(lib.eg_clip_adamw_step(self.handle(), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                              float(self._grad_norm or 0.0), lr, b1, b2, eps, wd, self._opt_step, st))

    def _dp_overlapped_step(self, mb, global_batch, moments):
        """One minibatch with the gradient sum of the actor + critic prefix (94 % of the bytes, and the whole clip-norm range)
        running on a side stream UNDER the encoders' backward: the NVLink traffic of the reduce phase is hidden, only the
        small tail reduce, AdamW on the slice and the parameter broadcast stay on the critical path."""
        import os
        dp, lib = self._dp, _lib.lib()
        W, rank, n_pad, n_ac = dp["world"], dp["rank"], dp["n_pad"], self.n_actor_critic
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device=self.dev)
            self._ev_a, self._ev_r = torch.cuda.Event(), torch.cuda.Event()
        main = torch.cuda.current_stream(self.dev)
        bar = dp["hdl"]["grads"].barrier
        mcg, mcp = C.c_void_p(dp["mc"]["grads"] or None), C.c_void_p(dp["mc"]["params"] or None)
        self.loss_backward(mb, global_batch, moments, part="mlp")
        self._ev_a.record(main)
        with torch.cuda.device(self.dev), torch.cuda.stream(self._comm):
            self._comm.wait_event(self._ev_a)
            bar(channel=1)                                         # every rank's actor / critic gradients are final
            _lib.check(lib.eg_dp_reduce_range(dp["ptrs"]["grads"], mcg, W, rank, n_pad, n_ac, 0, n_ac, 1, _lib.ptr(dp["gred"]),
                                              dp["ptrs"]["scratch"], _lib.ptr(dp["work"]), C.c_void_p(self._comm.cuda_stream)))
            self._ev_r.record(self._comm)
        self.loss_backward(None, part="encoders")
        lr, b1, b2, eps, wd = self._opt_hparams()
        self._opt_step += 1
        st = _lib.stream_ptr(self.dev)
        with torch.cuda.device(self.dev):
            main.wait_event(self._ev_r)
            bar(channel=0)                                         # every rank's encoder gradients are final
            _lib.check(lib.eg_dp_reduce_range(dp["ptrs"]["grads"], mcg, W, rank, n_pad, n_ac, n_ac, n_pad, 0, _lib.ptr(dp["gred"]),
                                              dp["ptrs"]["scratch"], _lib.ptr(dp["work"]), st))
            bar(channel=0)                                         # all slices read, all norm shares published
            _lib.check(lib.eg_dp_adamw_gather(dp["ptrs"]["params"], mcp, W, rank, n_pad, n_ac, _lib.ptr(dp["gred"]),
                                              _lib.ptr(dp["scratch"]), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                              float(self._grad_norm or 0.0), lr, b1, b2, eps, wd, self._opt_step, st))
            bar(channel=0)                                         # every rank's slice of the new parameters has landed

    def learn_minibatch(self, mb, global_batch: Optional[int] = None, moments=None, reduce_stats: bool = True):
        """One iteration of the inner loop of learn (:189-252). mb: Batch(obs, act, logp_old, adv, returns)
        with CUDA tensors. Returns a clone of the device stats tensor (rank-summed unless reduce_stats is False:
        learn() sums the statistics of all its minibatches in one collective at the end)."""
        import os
        # EG_DP_OVERLAP=1: gradient sum of the actor + critic prefix under the encoders' backward. Off by default: measured
        # SLOWER at N = 2 (13.1-13.7 vs 12.9 ms per iteration for 24..296 reduce CTAs) - the encoders' backward is only
        # ~0.15 ms long, and the extra barrier, the event hand-offs and the reduce CTAs it has to share the SMs with cost more
        if self._dp is not None and os.environ.get("EG_DP_OVERLAP", "0") == "1":
            self._dp_overlapped_step(mb, global_batch, moments)
        else:
            self.loss_backward(mb, global_batch, moments)
            self.optimizer_step()
        if reduce_stats and self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(self._stats, group=self.pg)
        return self._stats.clone()

    def learn(self, batch, batch_size: int, repeat: int, **kwargs):
        """learn (:182-265). batch: Batch of flat [N,...] CUDA tensors (obs dict, act, logp_old, adv, returns).
        batch_size is the PER-RANK minibatch size (global = batch_size * world). Minibatch order comes from
        np.random.permutation like tianshou's Batch.split(shuffle=True, merge_last=True)."""
        N = batch.act.shape[0]
        out = {k: [] for k in ("loss", "loss/clip", "loss/vf", "loss/ent", "loss/kld")}
        all_stats = []
        multi = self._world() > 1
        if multi:
            import torch.distributed as dist
        for step in range(repeat):
            perm = np.random.permutation(N)
            starts = list(range(0, N, batch_size))
            if len(starts) > 1 and N - starts[-1] < batch_size:
                starts = starts[:-1]                                   # merge_last
            bounds = [(s, N if i == len(starts) - 1 else s + batch_size) for i, s in enumerate(starts)]
            perm_dev = torch.as_tensor(perm, device=self.dev)          # one H2D copy of the permutation per repeat
            # ONE gather per field over the whole permutation (8 launches per repeat instead of 8 per minibatch); a
            # minibatch is a contiguous row range of the permuted copies
            take = perm_dev[:bounds[-1][1]] if bounds else perm_dev
            pobs = {k: v.index_select(0, take) for k, v in batch.obs.items()}
            pact, plogp = batch.act.index_select(0, take), batch.logp_old.index_select(0, take)
            padv, pret = batch.adv.index_select(0, take), batch.returns.index_select(0, take)
            mbs = [Batch(obs={k: v[s0:e0] for k, v in pobs.items()}, act=pact[s0:e0], logp_old=plogp[s0:e0],
                         adv=padv[s0:e0], returns=pret[s0:e0]) for s0, e0 in bounds]
            moms = None
            if self._norm_adv:                                         # advantage moments of every minibatch, ONE collective
                moms = torch.zeros(len(mbs), 3, dtype=torch.float64, device=self.dev)
                for i, mb in enumerate(mbs):
                    self.adv_moments(mb.adv.contiguous(), out=moms[i])
                if multi:
                    dist.all_reduce(moms, group=self.pg)
            rep_stats = [self.learn_minibatch(mb, moments=None if moms is None else moms[i], reduce_stats=False)
                         for i, mb in enumerate(mbs)]
            if rep_stats:
                S_rep = torch.stack(rep_stats)
                if multi:
                    dist.all_reduce(S_rep, group=self.pg)              # loss statistics of the repeat, ONE collective
                all_stats.extend(S_rep.unbind(0))
            if repeat > 1 and rep_stats and float(all_stats[-1][4].item()) >= 0.02:
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
    def _full_moments(self):
        """(exp_avg, exp_avg_sq) as full-length vectors: the sharded optimiser keeps only this rank's slice."""
        if self._dp is None:
            return self.exp_avg, self.exp_avg_sq
        import torch.distributed as dist
        full = [torch.empty(self._dp["n_pad"], dtype=torch.float32, device=self.dev) for _ in range(2)]
        dist.all_gather_into_tensor(full[0], self.exp_avg, group=self.pg)
        dist.all_gather_into_tensor(full[1], self.exp_avg_sq, group=self.pg)
        return full[0][:self.n_params], full[1][:self.n_params]

    def export_optim_state(self):
        exp_avg, exp_avg_sq = self._full_moments()
        for off, p in zip(self._offsets, self._ordered_params()):
            k = p.numel()
            self.optim.state[p] = {"step": torch.tensor(float(self._opt_step)),
                                   "exp_avg": exp_avg[off:off + k].view_as(p),
                                   "exp_avg_sq": exp_avg_sq[off:off + k].view_as(p)}
            off += k
        return self.optim.state_dict()
