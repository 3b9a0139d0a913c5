"""Device-resident rollout collection - the role tianshou's Collector + VectorReplayBuffer play in the
reference (main_ppo.py:178-183): policy(obs) under no-grad -> env.step -> store ->
finished envs reset immediately -> obs = obs_next, until n_step transitions are stored. Here the buffer is a set
of [T,E,...] CUDA tensors and the only host round trip per vector step is the `terminated` mask that decides
which envs to reset (tianshou needs the same information on the host).
"""
from __future__ import annotations

import numpy as np
import torch

from .ppo_policy import Batch


class RolloutBuffer:
    def __init__(self, T, E, dev):
        f = lambda *s: torch.zeros(T, E, *s, dtype=torch.float32, device=dev)
        self.T, self.E = T, E
        self.state, self.ego, self.dist, self.time = f(2, 402), f(2, 32), f(), f()
        self.act, self.logp, self.v_s, self.rew = f(128), f(), f(), f()
        self.term = torch.zeros(T, E, dtype=torch.uint8, device=dev)
        self.v_next = f()


class LazyStats(dict):
    """Collector statistics whose episode entries ('n/ep', 'rew', 'len', 'rews', 'lens') are read back from the device
    on first access, so a training loop that only logs every few collects never synchronises for them."""

    _LAZY = ("n/ep", "rew", "len", "rews", "lens")

    def _materialise(self):
        if "_term" in self:
            m = dict.pop(self, "_term").bool()
            r, l = dict.pop(self, "_ret_hist")[m], dict.pop(self, "_len_hist")[m]
            dict.__setitem__(self, "n/ep", int(r.numel()))
            col = dict.pop(self, "_collector", None)
            if col is not None:
                col.collect_episode += int(r.numel())
            if r.numel():
                dict.update(self, {"rew": float(r.mean()), "len": float(l.float().mean()), "rews": r.cpu().numpy(),
                                   "lens": l.cpu().numpy()})

    def __getitem__(self, k):
        if k in self._LAZY:
            self._materialise()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        if k in self._LAZY:
            self._materialise()
        return dict.get(self, k, default)

    def __contains__(self, k):
        if k in self._LAZY:
            self._materialise()
        return dict.__contains__(self, k)


class Collector:
    def __init__(self, policy, venv, host_boundary: bool = False):
        """host_boundary=True reproduces the reference's host-side data path (tianshou moves actions to numpy and
        the replay buffer lives in host memory): every vector step copies obs / act / reward / terminated through
        pinned host buffers. Used by bench.py's e2e measurement."""
        self.policy, self.venv = policy, venv
        self.dev, self.E = venv.dev, venv.E
        self.buf = None
        self.ep_ret = torch.zeros(self.E, device=self.dev)
        self.ep_len = torch.zeros(self.E, dtype=torch.int32, device=self.dev)
        self.host_boundary = host_boundary
        self.h2d_bytes = self.d2h_bytes = 0
        if host_boundary:
            pin = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt).pin_memory()
            E = self.E
            self._h = dict(state=pin(E, 2, 402), egosensing=pin(E, 2, 32), dist=pin(E, 1), time=pin(E, 1),
                           act=pin(E, 128), rew=pin(E), term=pin(E, dt=torch.uint8))
        self.collect_step = self.collect_episode = 0

    def reset(self):
        self.venv.reset()
        self.ep_ret.zero_(); self.ep_len.zero_()

    def _policy_obs(self):
        obs = self.venv.observation()
        if not self.host_boundary:
            return obs
        out = {}
        for k, v in obs.items():                       # D2H (env -> host buffer) then H2D (host -> policy)
            self._h[k].copy_(v, non_blocking=True)
            self.d2h_bytes += v.numel() * 4
        torch.cuda.current_stream(self.dev).synchronize()
        for k in obs:
            out[k] = self._h[k].to(self.dev, non_blocking=True)
            self.h2d_bytes += out[k].numel() * 4
        return out

    def _env_act(self, act):
        if not self.host_boundary:
            return act
        self._h["act"].copy_(act, non_blocking=True)   # tianshou: act -> numpy -> env (-> .cuda())
        torch.cuda.current_stream(self.dev).synchronize()
        self.d2h_bytes += act.numel() * 4
        self.h2d_bytes += act.numel() * 4
        return self._h["act"].to(self.dev, non_blocking=True)

    @torch.no_grad()
    def collect(self, n_step: int):
        """Collect n_step transitions (multiple of #envs). Returns (Batch for learn, stats dict)."""
        E = self.E
        assert n_step % E == 0, "n_step must be a multiple of the number of envs"
        T = n_step // E
        if self.buf is None or self.buf.T != T:
            self.buf = RolloutBuffer(T, E, self.dev)
        b, pol, venv = self.buf, self.policy, self.venv
        rets, lens = [], []
        maskable = hasattr(venv, "reset_masked") and getattr(venv, "sampler", None) is not None
        if not self.host_boundary and maskable:
            return self._collect_sync_free(T)
        if self.host_boundary and maskable:
            return self._collect_host(T)
        for t in range(T):
            obs = self._policy_obs()
            out = pol.forward(Batch(obs=obs), want_value=True)
            b.state[t].copy_(obs["state"]); b.ego[t].copy_(obs["egosensing"])
            b.dist[t].copy_(obs["dist"].view(-1)); b.time[t].copy_(obs["time"].view(-1))
            b.act[t].copy_(out.act); b.logp[t].copy_(out.logp); b.v_s[t].copy_(out.value)
            _, rew, term, _, _ = venv.step(self._env_act(out.act))
            b.rew[t].copy_(rew); b.term[t].copy_(term)
            self.ep_ret += rew; self.ep_len += 1
            if self.host_boundary:
                self._h["rew"].copy_(rew, non_blocking=True); self._h["term"].copy_(term, non_blocking=True)
                torch.cuda.current_stream(self.dev).synchronize()
                self.d2h_bytes += E * 5
                done = self._h["term"].nonzero().view(-1).to(self.dev)
            else:
                done = term.nonzero().view(-1)                      # the one host sync of the step
            if done.numel() > 0:
                rets.append(self.ep_ret[done].clone()); lens.append(self.ep_len[done].clone())
                self.ep_ret[done] = 0; self.ep_len[done] = 0
                venv.reset(done)
        # value bootstrap (ppo_policy.py:108-115): v_s_ of step t is the critic on obs_next = obs of step t+1
        _, v_last = pol.net_forward(venv.observation(), want_actor=False, want_critic=True)
        if T > 1:
            b.v_next[:-1].copy_(b.v_s[1:])
        b.v_next[-1].copy_(v_last)
        end = b.term.clone()
        end[-1] = 1                                                  # unfinished_index(): last stored step per env
        ret, adv = pol.compute_returns(b.v_s, b.v_next, b.rew, b.term, end)
        fl = lambda x: x.transpose(0, 1).reshape(T * E, *x.shape[2:]).contiguous()   # tianshou's env-major order
        batch = Batch(obs={"state": fl(b.state), "egosensing": fl(b.ego), "dist": fl(b.dist), "time": fl(b.time)},
                      act=fl(b.act), logp_old=fl(b.logp), adv=fl(adv), returns=fl(ret), v_s=fl(b.v_s))
        self.collect_step += n_step
        stats = {"n/st": n_step, "n/ep": 0}
        if rets:
            rl = torch.cat([torch.cat(rets), torch.cat(lens).to(torch.float32)]).cpu().numpy()      # ONE device -> host read
            r, l = rl[:len(rl) // 2], rl[len(rl) // 2:].astype(np.int32)
            self.collect_episode += r.size
            stats.update({"n/ep": int(r.size), "rew": float(r.mean()), "len": float(l.mean()), "rews": r, "lens": l})
        return batch, stats

    @torch.no_grad()
    def _collect_host(self, T: int):
        """collect() through the reference's host-side data path with TWO host synchronisations per vector step, the
        number tianshou's loop has (act -> numpy before env.step; obs_next / rew / done -> numpy after it): finished envs are
        restarted on the device from the step's `terminated` buffer, so the observation of the next step travels to the host
        together with this step's reward and flags."""
        E = self.E
        b, pol, venv = self.buf, self.policy, self.venv
        h, stream = self._h, torch.cuda.current_stream(self.dev)
        rets, lens = [], []

        def obs_to_host():
            for k, v in venv.observation().items():
                h[k].copy_(v, non_blocking=True)
                self.d2h_bytes += v.numel() * 4

        obs_to_host()
        stream.synchronize()
        if getattr(self, "_obs_dev", None) is None:                 # persistent device copies: no allocation per step
            self._obs_dev = {k: torch.empty_like(h[k], device=self.dev) for k in ("state", "egosensing", "dist", "time")}
        for t in range(T):
            obs = self._obs_dev
            for k in ("state", "egosensing", "dist", "time"):      # H2D (host observation -> policy)
                obs[k].copy_(h[k], non_blocking=True)
                self.h2d_bytes += obs[k].numel() * 4
            out = pol.forward(Batch(obs=obs), want_value=True)
            b.state[t].copy_(obs["state"]); b.ego[t].copy_(obs["egosensing"])
            b.dist[t].copy_(obs["dist"].view(-1)); b.time[t].copy_(obs["time"].view(-1))
            b.act[t].copy_(out.act); b.logp[t].copy_(out.logp); b.v_s[t].copy_(out.value)
            _, rew, term, _, _ = venv.step(self._env_act(out.act))  # sync 1: act -> host -> env
            b.rew[t].copy_(rew); b.term[t].copy_(term)
            self.ep_ret += rew; self.ep_len += 1
            venv.reset_masked(b.term[t])
            h["rew"].copy_(b.rew[t], non_blocking=True); h["term"].copy_(b.term[t], non_blocking=True)
            self.d2h_bytes += E * 5
            obs_to_host()
            stream.synchronize()                                    # sync 2: obs_next, rew, done -> host
            done = h["term"].nonzero().view(-1)
            if done.numel() > 0:
                done = done.to(self.dev)
                rets.append(self.ep_ret[done].clone()); lens.append(self.ep_len[done].clone())
                self.ep_ret[done] = 0; self.ep_len[done] = 0
        _, v_last = pol.net_forward(venv.observation(), want_actor=False, want_critic=True)
        if T > 1:
            b.v_next[:-1].copy_(b.v_s[1:])
        b.v_next[-1].copy_(v_last)
        end = b.term.clone()
        end[-1] = 1
        ret, adv = pol.compute_returns(b.v_s, b.v_next, b.rew, b.term, end)
        fl = lambda x: x.transpose(0, 1).reshape(T * E, *x.shape[2:]).contiguous()
        batch = Batch(obs={"state": fl(b.state), "egosensing": fl(b.ego), "dist": fl(b.dist), "time": fl(b.time)},
                      act=fl(b.act), logp_old=fl(b.logp), adv=fl(adv), returns=fl(ret), v_s=fl(b.v_s))
        self.collect_step += T * E
        stats = {"n/st": T * E, "n/ep": 0}
        if rets:
            rl = torch.cat([torch.cat(rets), torch.cat(lens).to(torch.float32)]).cpu().numpy()      # ONE device -> host read
            r, l = rl[:len(rl) // 2], rl[len(rl) // 2:].astype(np.int32)
            self.collect_episode += r.size
            stats.update({"n/ep": int(r.size), "rew": float(r.mean()), "len": float(l.mean()), "rews": r, "lens": l})
        return batch, stats

    @torch.no_grad()
    def _collect_sync_free(self, T: int):
        """Same transitions as collect(), with no host synchronisation inside the T vector steps: finished envs are
        restarted on the device from the step's `terminated` buffer (CrowdVectorEnv.reset_masked), episode statistics
        are snapshotted per step and read back once at the end."""
        E = self.E
        b, pol, venv = self.buf, self.policy, self.venv
        if getattr(self, "_ret_hist", None) is None or self._ret_hist.shape[0] != T:
            self._ret_hist = torch.zeros(T, E, device=self.dev)
            self._len_hist = torch.zeros(T, E, dtype=torch.int32, device=self.dev)
        for t in range(T):
            obs = venv.observation()
            out = pol.forward(Batch(obs=obs), want_value=True)
            b.state[t].copy_(obs["state"]); b.ego[t].copy_(obs["egosensing"])
            b.dist[t].copy_(obs["dist"].view(-1)); b.time[t].copy_(obs["time"].view(-1))
            b.act[t].copy_(out.act); b.logp[t].copy_(out.logp); b.v_s[t].copy_(out.value)
            _, rew, term, _, _ = venv.step(out.act)
            b.rew[t].copy_(rew); b.term[t].copy_(term)
            self.ep_ret += rew; self.ep_len += 1
            self._ret_hist[t].copy_(self.ep_ret); self._len_hist[t].copy_(self.ep_len)
            keep = (b.term[t] == 0)
            self.ep_ret *= keep; self.ep_len *= keep.to(torch.int32)
            venv.reset_masked(b.term[t])
        _, v_last = pol.net_forward(venv.observation(), want_actor=False, want_critic=True)
        if T > 1:
            b.v_next[:-1].copy_(b.v_s[1:])
        b.v_next[-1].copy_(v_last)
        end = b.term.clone()
        end[-1] = 1
        ret, adv = pol.compute_returns(b.v_s, b.v_next, b.rew, b.term, end)
        fl = lambda x: x.transpose(0, 1).reshape(T * E, *x.shape[2:]).contiguous()
        batch = Batch(obs={"state": fl(b.state), "egosensing": fl(b.ego), "dist": fl(b.dist), "time": fl(b.time)},
                      act=fl(b.act), logp_old=fl(b.logp), adv=fl(adv), returns=fl(ret), v_s=fl(b.v_s))
        self.collect_step += T * E
        stats = {"n/st": T * E, "n/ep": 0, "_term": b.term.clone(), "_ret_hist": self._ret_hist.clone(),
                 "_len_hist": self._len_hist.clone(), "_collector": self}
        return batch, LazyStats(stats)

    @torch.no_grad()
    def collect_episodes(self, n_episode: int, max_steps: int = 10000):
        """test_collector.collect(n_episode=...) (main_ppo.py:242). tianshou gives every env a quota (n_episode spread over
        the envs, the first n_episode % E envs one more) and stops stepping an env once its quota is met, so long episodes
        are not crowded out by envs that finish early; finished envs with quota left restart immediately.
        When the env has ``save_rollout`` set, every finished episode is written as a rollout pickle (crowd_env_2f.py:305-309)."""
        self.reset()
        E = self.E
        quota = np.full(E, n_episode // E, dtype=np.int64)
        quota[:n_episode % E] += 1
        done_count = np.zeros(E, dtype=np.int64)
        rets, lens, steps = [], [], 0
        save = bool(getattr(self.venv, "save_rollout", False))
        while int(done_count.sum()) < n_episode and steps < max_steps:
            out = self.policy.forward(Batch(obs=self.venv.observation()))
            pre = self.venv.begin_rollout_step() if save else None
            _, rew, term, _, _ = self.venv.step(out.act)
            active = torch.as_tensor(done_count < quota, device=self.dev)
            self.ep_ret += rew * active; self.ep_len += active.to(self.ep_len.dtype)
            if save:
                self.venv.record_rollout(pre, term & active.to(term.dtype))
            done = (term.bool() & active).nonzero().view(-1)
            if done.numel() > 0:
                rets.append(self.ep_ret[done].cpu().numpy()); lens.append(self.ep_len[done].cpu().numpy())
                self.ep_ret[done] = 0; self.ep_len[done] = 0
                done_count[done.cpu().numpy()] += 1
            fin = term.bool().nonzero().view(-1)                    # envs past their quota keep idling on fresh episodes
            if fin.numel() > 0:
                self.venv.reset(fin)
            steps += 1
        r = np.concatenate(rets)[:n_episode] if rets else np.zeros(0)
        l = np.concatenate(lens)[:n_episode] if lens else np.zeros(0)
        return {"n/ep": len(r), "rews": r, "lens": l, "rew": float(r.mean()) if len(r) else 0.0,
                "len": float(l.mean()) if len(l) else 0.0, "rew_std": float(r.std()) if len(r) else 0.0,
                "len_std": float(l.std()) if len(l) else 0.0}
