"""CPU harness around the oracle (TEST INFRASTRUCTURE): builds the oracle twin of the synthetic world that
egogen_b200.runtime.build_world creates on the GPU (same surrogate SMPL-X arrays, same seeded network weights,
same box scene) and runs it either batched or in the reference's execution shape - per-env sequential loop with
the 4x duplicated batch of crowd_env_2f.py:29-32,92 - for bench.py's cpu_baseline / --impl reference legs."""
import time

import numpy as np
import torch

from egogen_b200 import assets
from . import nets, ppo as oppo
from .env import CrowdEnvOracle
from .sdf import calc_sdf
from .smplx_lbs import SMPLXParserOracle


def build_oracle_world(seed=0, sdf_res=256, n_boxes=1):
    model = assets.make_surrogate_smplx(seed=0)
    markers = assets.marker_ids()
    combo = nets.ComboOracle()
    assets.fill_params_(combo.predictor, seed=11)
    assets.fill_params_(combo.regressor, seed=12, w_gain=0.7)
    with torch.no_grad():   # same human-scale damping as GAMMAPrimitiveComboGenOP.build_model
        combo.predictor.d_out.weight.mul_(0.02); combo.predictor.d_out.bias.mul_(0.02)
        combo.regressor.pnet.out_fc.weight.mul_(0.3)
    vp = assets.fill_params_(nets.VPoserEncoderOracle(), seed=31).eval()
    scene = assets.make_box_scene(seed, n_boxes=n_boxes)
    sdf = assets.rasterize_scene_sdf(scene, D=sdf_res)
    rings = assets.scene_polygon(scene)
    env = CrowdEnvOracle(SMPLXParserOracle(model, marker=markers), combo.eval(), vp, sdf, assets.rings_to_segments(rings),
                         markers, assets.feet_marker_idx(), assets.feet_vids())
    actor, critic, shared = nets.init_policy_nets(seed)
    return dict(env=env, sdf=sdf, actor=actor, critic=critic, shared=shared, model=model)


def sample_candidates_cpu(world, n, seed=0):
    """Upright 2-frame seeds at free positions of the scene (CPU analogue of BoxSceneSampler)."""
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(seed)
    env = world["env"]
    xb = torch.zeros(1, 93); xb[0, 3] = np.pi / 2
    out = env._lbs(xb, torch.zeros(1, 10))
    mask = torch.ones(out.vertices.shape[1], dtype=torch.bool); mask[assets.feet_vids()] = False
    z_lift = float(-out.vertices[0, mask, 2].min() + 0.08)
    pelvis_h = float(out.joints[0, 0, 2]) + z_lift

    def free_xy(k):
        acc = torch.empty(0, 2)
        while acc.shape[0] < k:
            xy = (torch.rand(4 * k + 64, 2, generator=g) * 2 - 1) * 3.2
            d = calc_sdf(torch.cat([xy, torch.full((xy.shape[0], 1), 0.9)], 1).unsqueeze(0), world["sdf"])[0]
            acc = torch.cat([acc, xy[d > 0.6]])
        return acc[:k]
    start, goal = free_xy(n), free_xy(n)
    yaw = torch.rand(n, generator=g) * 2 * np.pi
    R = np.zeros((n, 3, 3)); c, s = np.cos(yaw.numpy()), np.sin(yaw.numpy())
    R[:, 0, 0] = c; R[:, 0, 2] = s; R[:, 1, 0] = s; R[:, 1, 2] = -c; R[:, 2, 1] = 1
    aa = torch.as_tensor(Rotation.from_matrix(R).as_rotvec(), dtype=torch.float32)
    wp = torch.zeros(n, 2, 93)
    pose = torch.randn(n, 63, generator=g) * 0.05
    for t in range(2):
        wp[:, t, 0:2] = start; wp[:, t, 2] = z_lift; wp[:, t, 3:6] = aa; wp[:, t, 6:69] = pose
    goals = torch.cat([goal, torch.full((n, 1), pelvis_h)], 1)
    return wp, goals, torch.zeros(n, 10)


def init_env_state(world, n, seed=0):
    env = world["env"]
    wp, goals, betas = sample_candidates_cpu(world, n, seed)
    r = env.reset_from(wp, goals, betas)
    env.set_state(state=r["state"], seed=r["seed"], R0=r["R0"], T0=r["T0"], betas=betas, dist=r["dist"],
                  steps=torch.zeros(n, dtype=torch.int64), goal=goals)
    return dict(state=r["state"], egosensing=r["egosensing"], dist=r["obs_dist"].view(-1, 1), time=torch.ones(n, 1))


def _act(world, obs, g):
    with torch.no_grad():
        hx = world["shared"](obs)
        mu, lv = world["actor"](hx)
        val = world["critic"](hx).flatten()
        sig = torch.exp(lv.clamp(-2.5, 2.5)) ** 0.5
        act = mu + sig * torch.randn(mu.shape, generator=g)
        from torch.distributions import Independent, Normal
        logp = Independent(Normal(mu, sig), 1).log_prob(act)
    return act, logp, val


def run_iteration(world, n_envs, n_steps, reference_shaped, seed=0, n_minibatch=1):
    """One bounded PPO iteration on the CPU: n_steps vector steps of n_envs envs (policy forward + env step),
    GAE and one learn pass over the collected transitions. reference_shaped=True steps the envs one at a time
    with each env's batch replicated 4x (crowd_env_2f.py:92; DummyVectorEnv, main_ppo.py:97).
    Returns (seconds, env_steps)."""
    env = world["env"]
    g = torch.Generator().manual_seed(seed)
    obs = init_env_state(world, n_envs, seed)
    t0 = time.perf_counter()
    store = []
    for _ in range(n_steps):
        act, logp, val = _act(world, obs, g)
        if reference_shaped:
            full = {k: getattr(env, k).clone() for k in ("state", "seed", "R0", "T0", "betas", "dist", "steps", "goal")}
            outs = []
            for e in range(n_envs):
                env.set_state(**{k: v[e:e + 1].repeat(4, *([1] * (v.dim() - 1))) for k, v in full.items()})
                o = env.step(act[e:e + 1].repeat(4, 1))
                outs.append({k: o[k][:1] for k in ("state", "egosensing", "dist", "time", "reward", "terminated")})
                for k in full:
                    full[k][e] = getattr(env, k)[0]
            env.set_state(**full)
            o = {k: torch.cat([x[k] for x in outs]) for k in outs[0]}
        else:
            o = env.step(act)
        store.append(dict(obs=obs, act=act, logp=logp, val=val, rew=o["reward"], term=o["terminated"]))
        obs = dict(state=o["state"], egosensing=o["egosensing"], dist=o["dist"].view(-1, 1), time=o["time"].view(-1, 1))
    _, _, v_last = _act(world, obs, g)
    T = n_steps
    v_s = torch.stack([s["val"] for s in store]).t().reshape(-1).numpy()
    v_next = torch.stack([s["val"] for s in store][1:] + [v_last]).t().reshape(-1).numpy()
    rew = torch.stack([s["rew"] for s in store]).t().reshape(-1).numpy().astype(np.float64)
    term = torch.stack([s["term"] for s in store]).t().reshape(-1).numpy()
    unfinished = np.zeros(T * n_envs, bool); unfinished[T - 1::T] = True
    ret, adv = oppo.compute_episodic_return(v_s, v_next, rew, term, np.zeros_like(term), unfinished)
    cat = lambda k: torch.stack([s[k] for s in store]).transpose(0, 1).reshape(T * n_envs, *store[0][k].shape[1:])
    obs_all = {k: torch.stack([s["obs"][k] for s in store]).transpose(0, 1).reshape(T * n_envs, *store[0]["obs"][k].shape[1:])
               for k in store[0]["obs"]}
    act_all, logp_all = cat("act"), cat("logp")
    adv_t, ret_t = torch.as_tensor(adv, dtype=torch.float32), torch.as_tensor(ret, dtype=torch.float32)
    N = act_all.shape[0]
    perm = torch.as_tensor(np.random.default_rng(seed).permutation(N))
    for idx in torch.chunk(perm, max(1, min(n_minibatch, N))):          # learn(): one optimiser step per minibatch
        oppo.learn_minibatch(world["actor"], world["critic"], world["shared"], {k: v[idx] for k, v in obs_all.items()},
                             act_all[idx], logp_all[idx], adv_t[idx], ret_t[idx])
        oppo.clip_and_adamw(world["actor"], world["critic"], world["shared"])
    return time.perf_counter() - t0, n_envs * n_steps


# ---- bounded CPU samples of the secondary workloads (bench.py `secondary` entries) ----------------------------------------
def run_eval_steps(world, n_envs, n_steps, reference_shaped=True, seed=0):
    """Config 1 (main_ppo.py --watch): deterministic policy forward + env step, no learning. reference_shaped=True keeps the
    reference's 4x duplicated batch per env (crowd_env_2f.py:92). Returns (seconds, env_steps)."""
    env = world["env"]
    obs = init_env_state(world, n_envs, seed)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        with torch.no_grad():
            mu, _ = world["actor"](world["shared"](obs))
        if reference_shaped:
            full = {k: getattr(env, k).clone() for k in ("state", "seed", "R0", "T0", "betas", "dist", "steps", "goal")}
            outs = []
            for e in range(n_envs):
                env.set_state(**{k: v[e:e + 1].repeat(4, *([1] * (v.dim() - 1))) for k, v in full.items()})
                o = env.step(mu[e:e + 1].repeat(4, 1))
                outs.append({k: o[k][:1] for k in ("state", "egosensing", "dist", "time")})
                for k in full:
                    full[k][e] = getattr(env, k)[0]
            env.set_state(**full)
            o = {k: torch.cat([x[k] for x in outs]) for k in outs[0]}
        else:
            o = env.step(mu)
        obs = dict(state=o["state"], egosensing=o["egosensing"], dist=o["dist"].view(-1, 1), time=o["time"].view(-1, 1))
    return time.perf_counter() - t0, n_envs * n_steps


def run_crowd_steps(n_scenes=1, n_agents=4, n_steps=1, seed=0):
    """Config 4 (main_crowd_eval.py): `n_agents` CrowdEnv workers per scene stepped in DummyCrowdVectorEnv's order - the
    other agents' marker boxes are redistributed as holes before every worker's step (dummy_vector_env.py:33-39,78-82).
    Returns (seconds, agent_steps)."""
    model = assets.make_surrogate_smplx(seed=0)
    markers = assets.marker_ids()
    base = build_oracle_world(seed, sdf_res=64, n_boxes=0)
    floor = [np.asarray([[4.0, 4.0], [4.0, -4.0], [-4.0, -4.0], [-4.0, 4.0], [4.0, 4.0]], np.float64)]
    fl = np.asarray(floor[0], np.float32)
    tris = np.stack([fl[[0, 1, 2]], fl[[2, 3, 0]]])
    combo, vp = base["env"].combo, base["env"].vposer
    orcs = []
    for _ in range(n_agents):
        o = CrowdEnvOracle(SMPLXParserOracle(model, marker=markers), combo, vp, base["sdf"], assets.rings_to_segments(floor),
                           markers, assets.feet_marker_idx(), assets.feet_vids(), max_depth=60, box_mode=True, navmesh_tris=tris,
                           weight_look=0.1)
        o.crowd = True
        orcs.append(o)
    S, A = n_scenes, n_agents
    wp, goals, betas = sample_candidates_cpu(base, S * A, seed)
    for a in range(A):
        for sc in range(S):
            e = a * S + sc
            ang = 2 * np.pi * a / A
            pos = 3.0 * np.array([np.cos(ang), np.sin(ang)])
            d = wp[e, 1, :2] - wp[e, 0, :2]
            wp[e, 0, :2] = torch.as_tensor(pos, dtype=torch.float32)
            wp[e, 1, :2] = wp[e, 0, :2] + d
            goals[e, :2] = torch.as_tensor(-pos, dtype=torch.float32)
    sl = lambda a: slice(a * S, (a + 1) * S)
    bb, obs = [], []
    for a in range(A):
        orcs[a].holes = None
        bb.append(orcs[a].reset_from(wp[sl(a)], goals[sl(a)], betas[sl(a)])["bbox"])
    holes_for = lambda a: torch.stack([bb[o] for o in range(A) if o != a], dim=1)
    for a in range(A):
        orcs[a].holes = holes_for(a)
        r = orcs[a].reset_from(wp[sl(a)], goals[sl(a)], betas[sl(a)])
        orcs[a].set_state(state=r["state"], seed=r["seed"], R0=r["R0"], T0=r["T0"], betas=betas[sl(a)], dist=r["dist"],
                          steps=torch.zeros(S, dtype=torch.int64), goal=goals[sl(a)])
        obs.append(dict(state=r["state"], egosensing=r["egosensing"], dist=r["obs_dist"].view(-1, 1), time=torch.ones(S, 1)))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        for a in range(A):
            with torch.no_grad():
                mu, _ = base["actor"](base["shared"](obs[a]))
            orcs[a].holes = holes_for(a)
            o = orcs[a].step(mu)
            bb[a] = orcs[a].bbox
            obs[a] = dict(state=o["state"], egosensing=o["egosensing"], dist=o["dist"].view(-1, 1), time=o["time"].view(-1, 1))
    return time.perf_counter() - t0, S * A * n_steps


def run_cvae_train_steps(batch=64, n_frames=200, max_rollout=8, n_steps=1, seed=0):
    """Config 3 (GAMMAPrimitiveVAETrainOP.calc_loss_rollout + Adam): torch-CPU autograd over the oracle predictor on smooth
    synthetic sequences. Returns (seconds, primitives)."""
    from . import cvae_train
    pred = nets.PredictorOracle()
    assets.fill_params_(pred, seed=11)
    opt = torch.optim.Adam(pred.parameters(), lr=5e-4)
    g = torch.Generator().manual_seed(seed)
    base_ = torch.randn(batch, 1, 67, 3, generator=g) * 0.3
    walk = torch.cumsum(torch.randn(batch, n_frames, 1, 3, generator=g) * 0.01, dim=1)
    mk = (base_ + walk).reshape(batch, n_frames, 201).permute(1, 0, 2).contiguous()
    j = torch.randn(batch, 1, 22, 3, generator=g) * 0.3
    j[:, :, 1, 0] += 0.5; j[:, :, 2, 0] -= 0.5
    jts = (j + walk).permute(1, 0, 2, 3).contiguous()
    n_prim = 0
    t0 = time.perf_counter()
    for _ in range(n_steps):
        eps = [torch.randn(batch, 128, generator=g) for _ in range(max_rollout)]
        opt.zero_grad()
        loss = cvae_train.rollout_loss(pred, mk, jts, eps, max_rollout=max_rollout)
        loss.backward()
        opt.step()
        n_prim += batch * max_rollout
    return time.perf_counter() - t0, n_prim


def run_ego_depth_cpu(sdf, cam, H=64, W=64, fx=40.0, fy=40.0):
    """Config 5: the defining ego-depth oracle on a few agents. Returns (seconds, rays)."""
    from . import ego_depth as oed
    t0 = time.perf_counter()
    oed.ego_depth(sdf, cam, H, W, fx, fy)
    return time.perf_counter() - t0, cam.shape[0] * H * W
