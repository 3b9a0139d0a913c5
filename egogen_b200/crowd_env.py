"""Crowd environments - host-side mirror of the reference's motion/crowd_ppo/crowd_env_2f.py.

``CrowdVectorEnv`` steps E agents in ONE call of the CUDA library (eg_env_step): the reference runs
256 CrowdEnv instances sequentially under tianshou's DummyVectorEnv (main_ppo.py:97) and replicates
each agent 4x to dodge an smplx batch-size-1 bug (crowd_env_2f.py:29-32); here each agent is one row
of a device-resident batch and nothing leaves the GPU between the policy output and the next
observation. ``CrowdEnv`` keeps the reference's single-agent gymnasium surface on top of it
(``reset() -> (obs, {})``, ``step(z) -> (obs, float, bool, bool, {})``, ``seed``).

All arithmetic is in egogen_b200/csrc/env.cu; torch is used for buffers, RNG sampling and streams only.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import _lib, assets
from .sdf import calc_sdf, _grid3


# eg_env_set_scene registers conservative sign grids for the SDF it is given, keyed by the grid's device pointer
# (eg_sdf_prepare); several envs may share one grid tensor, so the registry entry is released with the last of them
_SDF_REFS = {}


def _sdf_unref(grid):
    if grid is None:
        return
    k = grid.data_ptr()
    n = _SDF_REFS.get(k, 0) - 1
    if n <= 0:
        _SDF_REFS.pop(k, None)
        _lib.lib().eg_sdf_release(_lib.ptr(grid))
    else:
        _SDF_REFS[k] = n


def default_cfg(finetuning: bool = False):
    """Values of crowd_ppo/cfg_samp20/MPVAEPolicy_samp_collision.yaml that the env reads."""
    return SimpleNamespace(
        modelconfig=SimpleNamespace(h_dim=512, z_dim=128, n_blocks=2, body_repr="ssm2_67_condi_marker_map",
                                    actfun="lrelu", min_logvar=-2.5, max_logvar=2.5, reproj_factor=0.5),
        lossconfig=SimpleNamespace(weight_vp=0.1, weight_floor=0.1, weight_skate=0.3, weight_target_dist=1.0,
                                   weight_face_target=0.1, weight_look_target=0.3, weight_success=0.5,
                                   weight_pene=0.1),
        trainconfig=SimpleNamespace(goal_thresh=0.1, max_depth=13, pene_thres=3, random_rotation_range=1),
        args=SimpleNamespace(gpu_index=0, random_seed=0))


def default_cfg_box():
    """MPVAEPolicy_samp_collision_2.yaml (box-scene env): look-target weight 0.1, max_depth 11, map 16 x 16 @ 0.8 m."""
    c = default_cfg()
    c.lossconfig.weight_look_target = 0.1
    c.lossconfig.pene_type = "body"
    c.trainconfig.max_depth = 11
    c.modelconfig.map_res, c.modelconfig.map_extent = 16, 0.8
    return c


def load_cfg(path: str):
    """Read the reference's policy yaml (OmegaConf in the reference, primitive_model.py:78-82)."""
    import yaml
    d = yaml.safe_load(open(path))
    ns = lambda x: SimpleNamespace(**{k: (ns(v) if isinstance(v, dict) else v) for k, v in x.items()})
    return ns(d)


class BoxSceneSampler:
    """Synthetic stand-in for exp_GAMMAPrimitive/utils/environments.py scene samplers:
    start pose / goal pairs in the free space of one rasterised scene, returned as world-frame 2-frame
    SMPL-X seeds. Gender is always 'male' and betas 0 like the reference samplers (environments.py:191,254)."""

    def __init__(self, scene_sdf: dict, lbs_model, device, floor_half: Optional[float] = None, seed: int = 0,
                 pose_noise: float = 0.05, min_goal_dist: float = 1.0):
        self.sdf, self.dev = scene_sdf, torch.device(device)
        # sampling square: the SDF volume's own x/y extent (grid coordinate p = (x - center) * scale spans [-1, 1]) unless
        # the caller gives a half-width; a scene that is not centred at the origin is therefore sampled where it is
        c = scene_sdf["center"].reshape(-1).to(torch.float32).cpu()
        sc = float(scene_sdf["scale"].reshape(-1)[0])
        self.cx, self.cy = float(c[0]), float(c[1])
        self.fh = float(floor_half) if floor_half is not None else min(4.0, 1.0 / max(sc, 1e-6))
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(seed)
        self.pose_noise, self.min_goal = pose_noise, min_goal_dist
        # upright offset: rest body rotated y-up -> z-up; lift so the lowest non-feet vertex clears the floor
        xb = torch.zeros(1, 93, device=self.dev)
        xb[0, 3] = np.pi / 2
        v, j, _ = lbs_model.forward(xb, torch.zeros(1, 10, device=self.dev), want_verts=True)
        mask = torch.ones(v.shape[1], dtype=torch.bool, device=self.dev)
        mask[assets.feet_vids()] = False
        self.z_lift = float(-v[0, mask, 2].min().item() + 0.08)
        self.pelvis_h = float(j[0, 0, 2].item()) + self.z_lift

    def seed(self, s: int):
        self.gen.manual_seed(int(s))
        self._pool = None

    def _free_xy(self, n):
        """n points whose column above the floor is at least 0.6 m from any obstacle (one host sync per call)."""
        out = torch.empty(0, 2, device=self.dev)
        ctr = torch.tensor([self.cx, self.cy], device=self.dev)
        for _ in range(64):                                       # bounded: a scene without free space must not hang
            if out.shape[0] >= n:
                return out[:n]
            xy = (torch.rand(2 * n + 64, 2, device=self.dev, generator=self.gen) * 2 - 1) * max(self.fh - 0.8, 0.1) + ctr
            pts = torch.cat([xy, torch.full((xy.shape[0], 1), 0.9, device=self.dev)], dim=1)
            d = calc_sdf(pts.unsqueeze(0), self.sdf)[0]
            out = torch.cat([out, xy[d > 0.6]], dim=0)
        if out.shape[0] >= n:
            return out[:n]
        raise _lib.EgError(f"BoxSceneSampler: found only {out.shape[0]} of {n} free start points within +-{self.fh - 0.8:.2f} m of "
                           f"({self.cx:.2f}, {self.cy:.2f}); pass floor_half or a sampler that knows the scene")

    def _refill(self, n):
        """Pre-generate a pool of n candidates entirely on the device so that next_body() is sync-free slicing."""
        dev = self.dev
        start, goal = self._free_xy(n), self._free_xy(n)
        for _ in range(8):
            bad = (goal - start).norm(dim=1) < self.min_goal
            nb = int(bad.sum())
            if nb == 0:
                break
            goal[bad] = self._free_xy(nb)
        yaw = torch.rand(n, device=dev, generator=self.gen) * (2 * np.pi)
        # global_orient = Rz(yaw) Rx(pi/2) as axis-angle via the quaternion product qz(yaw) * qx(pi/2)
        # (w = cos(yaw/2) / sqrt 2 stays away from +-1, so the log map is well conditioned)
        ch, sh, r = torch.cos(yaw / 2), torch.sin(yaw / 2), 0.5 ** 0.5
        q = torch.stack([ch * r, ch * r, sh * r, sh * r], dim=1)          # (w, x, y, z)
        ang = 2 * torch.acos(q[:, 0].clamp(-1, 1))
        aa = q[:, 1:] / torch.sin(ang / 2).unsqueeze(1) * ang.unsqueeze(1)
        cy, sy = torch.cos(yaw), torch.sin(yaw)
        wp = torch.zeros(n, 2, 93, device=dev)
        pose = torch.randn(n, 63, device=dev, generator=self.gen) * self.pose_noise
        fwd = torch.stack([sy, -cy], dim=1)                       # template forward (+z) after the rotation
        for t in range(2):
            wp[:, t, 0:2] = start + fwd * (0.02 * t)
            wp[:, t, 2] = self.z_lift
            wp[:, t, 3:6] = aa
            wp[:, t, 6:69] = pose
        goals = torch.cat([goal, torch.full((n, 1), self.pelvis_h, device=dev)], dim=1)
        self._pool = dict(world_params=wp, goals=goals, betas=torch.zeros(n, 10, device=dev))
        self._ptr = 0

    def next_body(self, n: int):
        if getattr(self, "_pool", None) is None or self._ptr + n > self._pool["goals"].shape[0]:
            self._refill(max(4096, 4 * n))
        a, b = self._ptr, self._ptr + n
        self._ptr = b
        return dict(world_params=self._pool["world_params"][a:b], goals=self._pool["goals"][a:b],
                    betas=self._pool["betas"][a:b], gender="male")


class CrowdVectorEnv:
    """E device-resident agents. Observation tensors are views of persistent CUDA buffers; copy them
    (``clone``) if they must survive the next ``step``."""

    def __init__(self, cfg, motion_model, lbs_model, vposer, scene_sdf: dict, scene_rings, sampler,
                 n_envs: int, device, feet_marker_idx=None, finetuning: bool = False,
                 capture_rollout: bool = False, debug_terms: bool = False, box_mode: bool = False, navmesh_tris=None):
        """box_mode=True selects the reference's crowd_env_2f_box.CrowdEnv semantics (2-D walkability-map penetration over
        ``navmesh_tris`` [F,3,2], penetration always terminates, weight_pene from the cfg); default is crowd_env_2f.CrowdEnv."""
        self.cfg, self.E, self.dev = cfg, int(n_envs), torch.device(device)
        self.box_mode = bool(box_mode)
        self.motion, self.lbs, self.vposer, self.sampler = motion_model, lbs_model, vposer, sampler
        self.finetuning = bool(finetuning)
        self.feet_marker_idx = list(feet_marker_idx if feet_marker_idx is not None else assets.feet_marker_idx())
        E, dev = self.E, self.dev
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.buf = dict(state=f(E, 2, 402), seed=f(E, 2, 93), R0=f(E, 3, 3), T0=f(E, 3), betas=f(E, 10), dist=f(E),
                        steps=torch.zeros(E, dtype=torch.int32, device=dev), goal=f(E, 3), ego=f(E, 2, 32),
                        obs_dist=f(E), obs_time=f(E), reward=f(E),
                        terminated=torch.zeros(E, dtype=torch.uint8, device=dev),
                        goal_reached=torch.zeros(E, dtype=torch.uint8, device=dev),
                        reward_terms=f(E, 8) if debug_terms else None,
                        out_markers=f(E, 20, 67, 3) if capture_rollout else None,
                        out_params=f(E, 20, 93) if capture_rollout else None,
                        out_pelvis=f(E, 20, 3) if capture_rollout else None)
        self._cbuf = _lib.EgEnvBuffers(**{k: (C.c_void_p(v.data_ptr()) if v is not None else None)
                                          for k, v in self.buf.items()})
        self._h = C.c_void_p()
        idx = self.dev.index or 0
        _lib.check(_lib.lib().eg_env_create(C.byref(self._config()), lbs_model._h, motion_model.handle(),
                                            vposer.handle(), idx, C.byref(self._h)))
        self.set_scene(scene_sdf, scene_rings)
        if self.box_mode:
            if navmesh_tris is None:
                raise _lib.EgError("box_mode needs navmesh_tris [F,3,2]")
            self._tris = torch.as_tensor(np.asarray(navmesh_tris), dtype=torch.float32, device=self.dev).contiguous()
            _lib.check(_lib.lib().eg_env_set_navmesh(self._h, _lib.ptr(self._tris), self._tris.shape[0]))
        self.action_dim = 128

    def _config(self):
        c, l, t = self.cfg.modelconfig, self.cfg.lossconfig, self.cfg.trainconfig
        w_pene = 0.1 if self.finetuning else 1.0                    # crowd_env_2f.py:268-271
        if self.box_mode:
            w_pene = float(l.weight_pene)                           # crowd_env_2f_box.py:226
        return _lib.EgEnvConfig(int(t.max_depth), int(self.finetuning), 40, (C.c_int32 * 6)(*self.feet_marker_idx),
                                float(c.reproj_factor), float(t.goal_thresh), float(l.weight_skate),
                                float(l.weight_floor), float(l.weight_face_target), float(l.weight_look_target),
                                float(l.weight_success), float(l.weight_target_dist), float(l.weight_vp),
                                w_pene, 7.0, int(self.box_mode), int(getattr(c, "map_res", 16)),
                                float(getattr(c, "map_extent", 0.8)), float(getattr(t, "pene_thres", 3)))

    def set_scene(self, scene_sdf, scene_rings):
        dev = self.dev
        old = getattr(self, "_grid", None)
        self.scene_sdf = scene_sdf
        self._grid = _grid3(scene_sdf)
        _SDF_REFS[self._grid.data_ptr()] = _SDF_REFS.get(self._grid.data_ptr(), 0) + 1
        _sdf_unref(old)
        self._center = scene_sdf["center"].to(torch.float32).reshape(-1).contiguous()
        self._scale = scene_sdf["scale"].to(torch.float32).reshape(-1).contiguous()
        skip = torch.zeros(assets.V_SMPLX, dtype=torch.uint8)
        skip[assets.feet_vids()] = 1
        self._skip = skip.to(dev)
        self._segs = torch.as_tensor(assets.rings_to_segments(scene_rings), dtype=torch.float64, device=dev).contiguous()
        g = self._grid
        _lib.check(_lib.lib().eg_env_set_scene(self._h, _lib.ptr(g), g.shape[0], g.shape[1], g.shape[2],
                                               _lib.ptr(self._center), _lib.ptr(self._scale), _lib.ptr(self._skip),
                                               _lib.ptr(self._segs), self._segs.shape[0]))

    def __len__(self):
        return self.E

    def seed(self, seed):
        self.sampler.seed(seed if isinstance(seed, int) else int(seed[0]))

    def observation(self):
        b = self.buf
        return {"state": b["state"], "egosensing": b["ego"], "dist": b["obs_dist"].view(-1, 1),
                "time": b["obs_time"].view(-1, 1)}

    def reset(self, env_ids: Optional[torch.Tensor] = None, max_tries: int = 50):
        """(Re)start the given envs (all by default): sample, canonicalise, reject starts that touch the scene
        (crowd_env_2f.py:326-396). One host sync per attempt (the accept mask), none on the step path."""
        dev = self.dev
        ids = torch.arange(self.E, dtype=torch.int32, device=dev) if env_ids is None else \
            torch.as_tensor(env_ids, dtype=torch.int32, device=dev).reshape(-1)
        tries = 0
        while ids.numel() > 0:
            if tries >= max_tries:
                raise _lib.EgError("reset: could not find collision-free start poses")
            n = ids.numel()
            s = self.sampler.next_body(n)
            accept = torch.zeros(n, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().eg_env_reset(self._h, C.byref(self._cbuf), _lib.ptr(ids.contiguous()), n,
                                                   _lib.ptr(s["world_params"].contiguous()),
                                                   _lib.ptr(s["goals"].contiguous()), _lib.ptr(s["betas"].contiguous()),
                                                   _lib.ptr(accept), _lib.stream_ptr(dev)))
            ids = ids[accept == 0]
            tries += 1
        return self.observation(), {}

    # ---- sync-free restart of finished episodes ------------------------------------------------------------
    _POOL_FIELDS = ("state", "seed", "R0", "T0", "betas", "dist", "steps", "goal", "ego", "obs_dist", "obs_time")

    def _refill_pool(self, pool: int = 2048, min_rows: int = 0):
        """(Re)build the pool of start candidates that eg_env_reset is known to accept, TOGETHER WITH the initial env state
        the reset computed for them: the sampler's candidates are run through the reset pipeline once on scratch slots (same
        accept test as reset(), crowd_env_2f.py:379-380 / the box env's map test; same canonicalisation, features and
        ego-sensing) and the accepted rows of every state buffer are kept. Restarting an episode is then a copy of one pool
        row - no SMPL-X pass, no accept mask to read back. One host sync per `pool` candidates."""
        dev = self.dev
        if getattr(self, "_scratch", None) is None:
            f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
            P = pool
            sb = dict(state=f(P, 2, 402), seed=f(P, 2, 93), R0=f(P, 3, 3), T0=f(P, 3), betas=f(P, 10), dist=f(P),
                      steps=torch.zeros(P, dtype=torch.int32, device=dev), goal=f(P, 3), ego=f(P, 2, 32),
                      obs_dist=f(P), obs_time=f(P), reward=f(P), terminated=torch.zeros(P, dtype=torch.uint8, device=dev),
                      goal_reached=None, reward_terms=None, out_markers=None, out_params=None, out_pelvis=None)
            self._scratch = sb
            self._scratch_c = _lib.EgEnvBuffers(**{k: (C.c_void_p(v.data_ptr()) if v is not None else None)
                                                   for k, v in sb.items()})
            self._scratch_ids = torch.arange(P, dtype=torch.int32, device=dev)
            self._cursor = torch.zeros(2, dtype=torch.int64, device=dev)      # {int64 cursor, uint32 ticket + pad}
            self._cursor_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        parts, have = [], 0
        while have < max(min_rows, pool // 2):
            s = self.sampler.next_body(pool)
            accept = torch.zeros(pool, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().eg_env_reset(self._h, C.byref(self._scratch_c), _lib.ptr(self._scratch_ids), pool,
                                                   _lib.ptr(s["world_params"].contiguous()), _lib.ptr(s["goals"].contiguous()),
                                                   _lib.ptr(s["betas"].contiguous()), _lib.ptr(accept), _lib.stream_ptr(dev)))
            ok = accept != 0                                         # the one host sync of the refill
            part = {k: self._scratch[k][ok].clone() for k in self._POOL_FIELDS}
            part["world_params"] = s["world_params"][ok].clone()
            parts.append(part)
            have += int(ok.sum())
        self._vpool = {k: torch.cat([p_[k] for p_ in parts]).contiguous() for k in parts[0]}
        self._pool_rows = int(self._vpool["goal"].shape[0])
        bufs = {k: None for k in self.buf}
        bufs.update({k: self._vpool[k] for k in self._POOL_FIELDS})
        self._pool_c = _lib.EgEnvBuffers(**{k: (C.c_void_p(v.data_ptr()) if v is not None else None) for k, v in bufs.items()})
        self._cursor.zero_()
        self._used_bound, self._cursor_ev = 0, None
        self.pool_refills = getattr(self, "pool_refills", 0) + 1

    def _pool_reserve(self, n_max: int):
        """Make sure the next restart (which consumes at most n_max rows) cannot run past the pool. The exact consumption
        lives in the device cursor; the host keeps an upper bound (n_max per call since the last value it has SEEN) and
        tightens it from an asynchronous cursor snapshot when one has landed, or with one blocking read when the bound
        says the pool might be exhausted. Only a pool that is really used up is rebuilt."""
        if getattr(self, "_vpool", None) is None:
            self._refill_pool(min_rows=n_max)
        if self._cursor_ev is not None and self._cursor_ev[0].query():
            ev, bound_then = self._cursor_ev                         # true cursor after the call whose bound was bound_then
            self._used_bound = int(self._cursor_host[0]) + (self._used_bound - bound_then)
            self._cursor_ev = None
        if self._used_bound + n_max > self._pool_rows:
            self._used_bound = int(self._cursor[0].item())           # blocking read: rare (see docstring)
            self._cursor_ev = None
            if self._used_bound + n_max > self._pool_rows:
                self._refill_pool(min_rows=n_max)

    def _validated_candidates(self, n: int):
        """n pre-validated start candidates (pool rows at the cursor, which advances by n): tests / direct resets."""
        self._pool_reserve(n)
        a = int(self._cursor[0].item())
        self._cursor[0] += n
        self._used_bound, self._cursor_ev = a + n, None
        out = {k: v[a:a + n] for k, v in self._vpool.items()}
        out["goals"] = out["goal"]
        return out

    def reset_masked(self, mask: torch.Tensor):
        """Restart the envs whose mask entry is non-zero (device uint8 / bool [E], e.g. the `terminated` buffer) without
        any host synchronisation: the k-th flagged env takes the k-th unused row of the pre-validated pool
        (eg_env_restart_from_pool: device-side rank + cursor), so only finished episodes consume candidates."""
        self._pool_reserve(self.E)
        m = mask if mask.dtype == torch.uint8 else mask.to(torch.uint8)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_env_restart_from_pool(C.byref(self._cbuf), C.byref(self._pool_c), self._pool_rows,
                                                           _lib.ptr(m.contiguous()), self.E, _lib.ptr(self._cursor),
                                                           _lib.stream_ptr(self.dev)))
        self._used_bound += self.E
        if self._cursor_ev is None:                                  # asynchronous snapshot of the true consumption
            self._cursor_host.copy_(self._cursor[:1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))
            self._cursor_ev = (ev, self._used_bound)
        return m

    def reset_from(self, env_ids, world_params, goals, betas):
        """Deterministic reset from explicit candidates (tests / parity); returns the accept mask."""
        dev = self.dev
        ids = torch.as_tensor(env_ids, dtype=torch.int32, device=dev).contiguous()
        accept = torch.zeros(ids.numel(), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().eg_env_reset(self._h, C.byref(self._cbuf), _lib.ptr(ids), ids.numel(),
                                               _lib.ptr(_lib.f32c(world_params, dev)), _lib.ptr(_lib.f32c(goals, dev)),
                                               _lib.ptr(_lib.f32c(betas, dev)), _lib.ptr(accept), _lib.stream_ptr(dev)))
        return accept

    _STATE_KEYS = ("state", "seed", "R0", "T0", "betas", "dist", "steps", "goal", "ego", "obs_dist", "obs_time", "reward",
                   "terminated", "goal_reached")

    def _ids(self, id):
        if id is None:
            return None
        ids = torch.as_tensor([id] if isinstance(id, (int, np.integer)) else np.asarray(id), dtype=torch.long, device=self.dev)
        return None if (ids.numel() == self.E and bool((ids == torch.arange(self.E, device=self.dev)).all())) else ids

    def step(self, action_z: torch.Tensor, id=None):
        """action_z [E,128] CUDA float32 -> (obs, reward [E], terminated [E] uint8, truncated [E], info).
        ``id`` (tianshou BaseVectorEnv.step, dummy_vector_env.py:41-45): step only those envs - actions, returned tensors and
        ``info`` then follow the order of ``id``; the other envs keep their state. info[i]['env_id'] names the env of row i."""
        ids = self._ids(id)
        z = _lib.f32c(action_z, self.dev)
        keep = None
        if ids is not None:
            if z.shape != (ids.numel(), 128):
                raise _lib.EgError(f"action must be [{ids.numel()},128] for {ids.numel()} env ids")
            full = torch.zeros(self.E, 128, device=self.dev)
            full[ids] = z
            z = full
            keep = {k: self.buf[k].clone() for k in self._STATE_KEYS if self.buf.get(k) is not None}
        if z.shape != (self.E, 128):
            raise _lib.EgError(f"action must be [{self.E},128]")
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().eg_env_step(self._h, C.byref(self._cbuf), _lib.ptr(z), self.E,
                                              _lib.stream_ptr(self.dev)))
        b = self.buf
        if ids is None:
            return self.observation(), b["reward"], b["terminated"], torch.zeros_like(b["terminated"]), \
                [{"env_id": j} for j in range(self.E)]
        mask = torch.zeros(self.E, dtype=torch.bool, device=self.dev)
        mask[ids] = True
        out = {k: b[k][ids].clone() for k in ("state", "ego", "obs_dist", "obs_time", "reward", "terminated")}
        for k, v in keep.items():                                  # envs that were not stepped keep their state
            mk = mask.view(-1, *([1] * (v.dim() - 1)))
            b[k].copy_(torch.where(mk, b[k], v))
        obs = {"state": out["state"], "egosensing": out["ego"], "dist": out["obs_dist"].view(-1, 1), "time": out["obs_time"].view(-1, 1)}
        return obs, out["reward"], out["terminated"], torch.zeros_like(out["terminated"]), [{"env_id": int(j)} for j in ids.tolist()]

    # ---- rollout pickles at episode end (crowd_env_2f.py:154-155,305-309 -> utils.save_rollout_results) -----------
    save_rollout = False
    rollout_dir = "./log/eval_results/"

    def begin_rollout_step(self):
        """Call BEFORE step() when save_rollout is on: the canonical frame a primitive is recorded with is the one it was
        generated in (:155 precedes the re-canonicalisation at :247)."""
        return self.buf["R0"].clone(), self.buf["T0"].clone()

    def record_rollout(self, pre, terminated):
        """Call AFTER step() with begin_rollout_step()'s result: appends every env's primitive to its episode record and
        writes ``motion_<time>.pkl`` for the envs whose episode ended. Host-side (one D2H of the flags); evaluation only.
        Returns the files written."""
        if self.buf.get("out_markers") is None:
            raise _lib.EgError("record_rollout needs capture_rollout=True")
        from .utils import save_rollout_results
        if getattr(self, "_outmps", None) is None:
            self._outmps = [[] for _ in range(self.E)]
            self._wpath0 = [None] * self.E
        R0, T0 = pre
        b = self.buf
        term = terminated.cpu().numpy().astype(bool)
        files = []
        for e in range(self.E):
            if not self._outmps[e]:
                self._wpath0[e] = T0[e].clone()                      # pelvis of the start frame = origin of the first frame
            self._outmps[e].append([b["out_markers"][e:e + 1].clone(), b["out_params"][e:e + 1].clone(), b["betas"][e].clone(),
                                    "male", R0[e], T0[e].view(1, 3), b["out_pelvis"][e:e + 1].clone(), "2-frame"])
            if term[e]:
                scene = {"wpath": torch.stack([self._wpath0[e], b["goal"][e]]),
                         "navmesh_path": getattr(self.sampler, "navmesh_path", None),
                         "scene_path": getattr(self.sampler, "scene_path", None)}
                files.append(save_rollout_results(scene, self._outmps[e], self.rollout_dir))
                self._outmps[e] = []
        return files

    def get_env_attr(self, key: str, id=None):
        """tianshou BaseVectorEnv.get_env_attr: one value per env (a row of the device buffer `key`, or the Python attribute
        shared by all envs)."""
        ids = self._ids(id)
        rows = range(self.E) if ids is None else ids.tolist()
        src = self.buf.get(key) if key in self.buf else getattr(self, key)
        if torch.is_tensor(src) and src.dim() >= 1 and src.shape[0] == self.E:
            return [src[j] for j in rows]
        return [src for _ in rows]

    def set_env_attr(self, key: str, value, id=None):
        """tianshou BaseVectorEnv.set_env_attr: write `value` (one value, or one per env) into the rows of buffer `key`."""
        ids = self._ids(id)
        rows = list(range(self.E)) if ids is None else ids.tolist()
        dst = self.buf.get(key) if key in self.buf else getattr(self, key, None)
        if not (torch.is_tensor(dst) and dst.dim() >= 1 and dst.shape[0] == self.E):
            setattr(self, key, value)
            return
        vals = value if isinstance(value, (list, tuple)) and len(value) == len(rows) else [value] * len(rows)
        for j, v in zip(rows, vals):
            dst[j] = torch.as_tensor(v, dtype=dst.dtype, device=self.dev).reshape(dst[j].shape)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().eg_env_destroy(self._h)
            self._h = None
            _sdf_unref(getattr(self, "_grid", None))            # drop the conservative sign grids registered for this SDF
            self._grid = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def calc_egosensing(joints_local, R0, T0, scene_rings, ray_len: float = 7.0, holes=None):
    """``CrowdEnv._calc_egosensing`` (crowd_env_2f.py:524-613) as a stand-alone operator: joints_local [n,2,127,3] CUDA
    float32 body-frame joints of the 2-frame seed, R0 [n,3,3], T0 [n,3], the scene polygon rings (exterior + holes) and
    optional per-item hole rectangles [n,H,4] -> [n,2,32] ray distances mapped to [-1,1]."""
    dev = joints_local.device
    j = _lib.f32c(joints_local, dev)
    n = j.shape[0]
    segs = torch.as_tensor(assets.rings_to_segments(scene_rings), dtype=torch.float64, device=dev).contiguous()
    out = torch.zeros(n, 2, 32, dtype=torch.float32, device=dev)
    hl = None if holes is None else _lib.f32c(holes, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().eg_egosensing(_lib.ptr(j), _lib.ptr(_lib.f32c(R0, dev)), _lib.ptr(_lib.f32c(T0, dev)), n,
                                            _lib.ptr(segs), segs.shape[0], float(ray_len), _lib.ptr(hl),
                                            0 if hl is None else hl.shape[1], _lib.ptr(out), _lib.stream_ptr(dev)))
    return out


class CrowdSceneVectorEnv(CrowdVectorEnv):
    """Multi-agent crowd dynamics: ``n_scenes`` scenes x ``n_agents`` agents, the batched form of
    ``DummyCrowdVectorEnv([CrowdEnv(init_env), CrowdEnv(init_env2), ...])`` (motion/crowd_ppo/dummy_vector_env.py:29-128,
    main_crowd_eval.py:47) over ``crowd_env_crowd_eval.CrowdEnv`` (box-env arithmetic + the other agents' marker
    bounding boxes as holes of the floor polygon, :66-75,345-352,796-822; termination on goal / max_depth only, :367;
    fixed start data, no rejection, :391-405).

    Env index = agent * n_scenes + scene (agent-major), so the agents that step together are one contiguous slice.
    ``sequential=True`` reproduces the reference's update order exactly: ``update_holes_for_each_agent()`` runs before
    every worker's step (:78-82) and DummyVectorEnv workers step synchronously, so agent a already sees the NEW
    boxes of agents < a of the same vector step. ``sequential=False`` steps all agents against the boxes of the
    previous vector step in one launch sequence (4x fewer launches; a documented deviation)."""

    FLOOR = [[4.0, 4.0], [4.0, -4.0], [-4.0, -4.0], [-4.0, 4.0], [4.0, 4.0]]       # crowd_env_crowd_eval.py:398

    def __init__(self, cfg, motion_model, lbs_model, vposer, scene_sdf: dict, n_scenes: int, device, n_agents: int = 4,
                 sequential: bool = True, feet_marker_idx=None, debug_terms: bool = False, capture_rollout: bool = False):
        self.S, self.A, self.sequential = int(n_scenes), int(n_agents), bool(sequential)
        fl = np.asarray(self.FLOOR, np.float32)
        tris = np.stack([fl[[0, 1, 2]], fl[[2, 3, 0]]])                              # floor square as two triangles
        super().__init__(cfg, motion_model, lbs_model, vposer, scene_sdf, [np.asarray(self.FLOOR, np.float64)], None,
                         self.S * self.A, device, feet_marker_idx, False, capture_rollout, debug_terms,
                         box_mode=True, navmesh_tris=tris)
        E, dev = self.E, self.dev
        self.bbox = torch.zeros(E, 4, dtype=torch.float32, device=dev)
        self.holes = torch.zeros(E, max(self.A - 1, 1), 4, dtype=torch.float32, device=dev)
        # one EgEnvBuffers view per agent slot (rows [a*S, (a+1)*S) of every buffer)
        self._slices = []
        for a in range(self.A):
            sl = slice(a * self.S, (a + 1) * self.S)
            self._slices.append(_lib.EgEnvBuffers(**{k: (C.c_void_p(v[sl].data_ptr()) if v is not None else None)
                                                     for k, v in self.buf.items()}))

    def _set_crowd(self, a=None, with_holes=True):
        lo = 0 if a is None else a * self.S
        n_h = self.A - 1 if (with_holes and self.A > 1) else 0
        _lib.check(_lib.lib().eg_env_set_crowd(self._h, _lib.ptr(self.holes[lo:]) if n_h else None, n_h,
                                               _lib.ptr(self.bbox[lo:]), 0))

    def update_holes_for_each_agent(self):
        """dummy_vector_env.py:33-39: holes of agent a = the current boxes of every other agent of its scene (ascending)."""
        if self.A < 2:
            return
        bb = self.bbox.view(self.A, self.S, 4)
        hv = self.holes.view(self.A, self.S, self.A - 1, 4)
        for a in range(self.A):
            others = [o for o in range(self.A) if o != a]
            hv[a].copy_(bb[others].permute(1, 0, 2))

    def reset_from(self, env_ids, world_params, goals, betas):
        """Start data is fixed per agent (crowd_env_crowd_eval.py:55-78): first pass computes every agent's box
        (CrowdEnv.__init__), then the holes are distributed (DummyCrowdVectorEnv.__init__) and the observation is
        built against them (reset -> _get_feature / _calc_egosensing)."""
        self._set_crowd(None, with_holes=False)
        super().reset_from(env_ids, world_params, goals, betas)
        self.update_holes_for_each_agent()
        self._set_crowd(None, with_holes=True)
        return super().reset_from(env_ids, world_params, goals, betas)

    def reset(self, env_ids=None, max_tries: int = 50):
        raise _lib.EgError("CrowdSceneVectorEnv starts from explicit per-agent data: use reset_from (main_crowd_eval.py:47)")

    def step(self, action_z: torch.Tensor, id=None):
        if self._ids(id) is not None:
            raise NotImplementedError("the crowd scene steps all agents of all scenes together (DummyCrowdVectorEnv.step with id=None)")
        z = _lib.f32c(action_z, self.dev)
        if z.shape != (self.E, 128):
            raise _lib.EgError(f"action must be [{self.E},128]")
        lib = _lib.lib()
        with torch.cuda.device(self.dev):
            if self.sequential:
                for a in range(self.A):
                    self.update_holes_for_each_agent()              # before every worker's step (:78-82)
                    self._set_crowd(a)
                    _lib.check(lib.eg_env_step(self._h, C.byref(self._slices[a]), _lib.ptr(z[a * self.S:]), self.S,
                                               _lib.stream_ptr(self.dev)))
            else:
                self.update_holes_for_each_agent()
                self._set_crowd(None)
                _lib.check(lib.eg_env_step(self._h, C.byref(self._cbuf), _lib.ptr(z), self.E, _lib.stream_ptr(self.dev)))
        b = self.buf
        return self.observation(), b["reward"], b["terminated"], torch.zeros_like(b["terminated"]), \
            [{"env_id": j} for j in range(self.E)]


class CrowdEnv:
    """Single-agent environment with the reference constructor and gymnasium surface
    (crowd_env_2f.py:34-51,78,320,519). ``init_env`` is the reference's 12-element list
    [cfg, genop_male, genop_female, bm_path, scene_sampler, parser_1f, parser_2f, parser_mp,
     feet_marker_idx, marker_ids, vposer, scene_sdf]; the scene polygon comes from
    ``scene_sampler.scene_rings`` (the reference reads it from the sampler dict's 'shapely_poly')."""

    def __init__(self, init_env, save_rollout=True, render=False, finetuning=False):
        if render:
            raise NotImplementedError("pyrender visualisation is out of scope (SURVEY.md section 2 #17)")
        (self.cfg, genop_male, _genop_female, _bm_path, self.scene_sampler, _p1, _p2, parser_mp,
         self.feet_marker_idx, self.marker, self.vposer, self.scene_sdf) = init_env
        self.save_rollout, self.finetuning = save_rollout, finetuning
        dev = parser_mp.device
        self.action_space = SimpleNamespace(low=-6.0, high=6.0, shape=(128,))
        self.observation_space = {"state": (2, 402), "egosensing": (2, 32), "dist": (1,), "time": (1,)}
        self._venv = CrowdVectorEnv(self.cfg, genop_male.model, parser_mp.bm_male, self.vposer, self.scene_sdf,
                                    self.scene_sampler.scene_rings, self.scene_sampler, 1, dev,
                                    self.feet_marker_idx, finetuning, capture_rollout=save_rollout)
        self.outmps, self.flag, self.steps = [], False, 0
        self.rollout_dir = "./log/eval_results/"                   # crowd_env_2f.py:309
        self.body_scene_data = None

    def seed(self, seed):
        np.random.seed(seed)
        torch.manual_seed(seed)
        self._venv.seed(seed)

    @staticmethod
    def _single(obs):
        return {"state": obs["state"][0], "egosensing": obs["egosensing"][0], "dist": obs["dist"][0],
                "time": obs["time"][0]}

    def reset(self, seed=None, options=None):
        self.flag, self.steps, self.outmps = False, 0, []
        obs, info = self._venv.reset()
        b = self._venv.buf
        # what save_rollout_results reads from the sampler dict (utils.py:10-28): the way-points (start pelvis, goal) and
        # the scene / navmesh file names when the sampler knows them
        self.body_scene_data = {"wpath": torch.stack([b["T0"][0], b["goal"][0]]).clone(),
                                "navmesh_path": getattr(self.scene_sampler, "navmesh_path", None),
                                "scene_path": getattr(self.scene_sampler, "scene_path", None)}
        return self._single(obs), info

    def step(self, action_z):
        if self.flag:
            raise RuntimeError("the episode should be terminated! do not collect undefined states")
        z = torch.as_tensor(action_z, dtype=torch.float32, device=self._venv.dev).reshape(1, 128)
        b = self._venv.buf
        R0, T0 = b["R0"][0].clone(), b["T0"][0].clone().view(1, 3)      # frame of THIS primitive (:155 precedes :247)
        obs, rew, term, _, _ = self._venv.step(z)
        self.steps += 1
        terminated = bool(term[0].item())
        if self.save_rollout:
            self.outmps.append([b["out_markers"].clone(), b["out_params"].clone(), b["betas"][0].clone(), "male",
                                R0, T0, b["out_pelvis"].clone(), "2-frame"])
            if terminated:
                self.flag = True
                from .utils import save_rollout_results
                self.last_rollout_file = save_rollout_results(self.body_scene_data, self.outmps, self.rollout_dir)   # :305-309
        return self._single(obs), float(rew[0].item()), terminated, False, {}
