"""Pin the CPU oracle against golden vectors produced by the reference's own classes
(tests/golden/gen_golden.py). CPU-only."""
import os

import numpy as np
import torch

from egogen_b200.assets import fill_params_
from oracle import nets, sdf as osdf, tgm
from oracle.smplx_lbs import SMPLXParserOracle


def _t(a):
    return torch.as_tensor(np.asarray(a))


def test_calc_sdf_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "sdf_golden.npz"))
    for k in "ab":
        d = {"center": _t(g[f"center_{k}"]), "scale": _t(g[f"scale_{k}"]), "sdf": _t(g[f"grid_{k}"])}
        out = osdf.calc_sdf(_t(g[f"pts_{k}"]), d)
        assert torch.equal(out, _t(g[f"out_{k}"]))          # same ATen op => bit-identical
        out2, idx = osdf.calc_sdf_explicit(_t(g[f"pts_{k}"]), d)
        assert torch.allclose(out2, out, atol=2e-6, rtol=0)
        assert torch.equal(out2 < 0, out < 0) or ((out2 < 0) != (out < 0)).sum() <= 1
        D = torch.tensor(d["sdf"].shape)
        assert (idx >= 0).all() and (idx <= D - 1).all()


def test_predictor_regressor_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "nets_motion_golden.npz"))
    pred = fill_params_(nets.PredictorOracle().eval(), seed=11)
    reg = fill_params_(nets.RegressorOracle().eval(), seed=12, w_gain=0.7)
    with torch.no_grad():
        Y = pred.sample_prior(_t(g["X"]), _t(g["z"]))
        xb = reg.forward_cont(Y.reshape(-1, 201), _t(g["betas"]))
        rot = tgm.cont2rotmat(xb[:, 3:135].contiguous().view(xb.shape[0], -1, 6))
    assert torch.allclose(Y, _t(g["Y"]), atol=1e-6, rtol=1e-6)
    assert torch.allclose(xb, _t(g["xb_cont"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(rot, _t(g["rotmat"]), atol=1e-5, rtol=1e-5)


def test_policy_nets_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "nets_policy_golden.npz"))
    actor = fill_params_(nets.ActorOracle(), seed=21)
    critic = fill_params_(nets.CriticOracle(), seed=22)
    shared = fill_params_(nets.PolicyBaseOracle(), seed=23)
    obs = {k: _t(g[k]) for k in ("state", "egosensing", "dist", "time")}
    with torch.no_grad():
        hx = shared(obs)
        mu, logvar = actor(hx)
        val = critic(hx)
    assert torch.allclose(hx, _t(g["hx"]), atol=1e-6, rtol=1e-6)
    assert torch.allclose(mu, _t(g["mu"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(logvar, _t(g["logvar"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(val, _t(g["value"]), atol=1e-5, rtol=1e-5)


def test_policy_param_counts():
    a, c, s = nets.init_policy_nets(0)
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert (n(a), n(c), n(s)) == (5608192, 5314177, 2245632)
    p, r = nets.PredictorOracle(), nets.RegressorOracle()
    assert (n(p), n(r)) == (2455497, 398239)


def test_coordinate_extractor_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "coord_golden.npz"))
    R, T = SMPLXParserOracle.new_coordinate_from_joints(_t(g["jts"]))
    assert torch.allclose(R, _t(g["R"]), atol=1e-7) and torch.equal(T, _t(g["T"]))
    R2, T2 = SMPLXParserOracle.new_coordinate_from_joints(_t(g["jts_loco"]))
    assert torch.allclose(R2, _t(g["R_loco"]), atol=1e-7)
    # model-free invariant of the reference fixture subseq_00343.npz:
    assert torch.allclose(R2[0], torch.eye(3), atol=1e-3) and T2.abs().max() < 1e-3


# ---- training objectives: the reference's own train ops (tests/golden/gen_train_golden.py) -----------------------------
def _gradnorms(module):
    return np.array([p.grad.norm().item() if p.grad is not None else -1.0 for p in module.parameters()])


def _regressor_weights(reg, seed):
    from egogen_b200.assets import fill_params_
    fill_params_(reg, seed=seed, w_gain=0.5)
    with torch.no_grad():
        gg = torch.Generator().manual_seed(seed)
        reg.pnet.out_fc.bias[3:135] = torch.tensor([1.0, 0.0, 0.0, 1.0, 0.0, 0.0]).repeat(22) + torch.randn(132, generator=gg) * 0.3


def _close(a, b, rtol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-6) + 1e-9)


def test_cvae_training_oracle_matches_reference_trainop(golden_dir):
    """oracle.cvae_train.primitive_loss / rollout_loss vs GAMMAPrimitiveVAETrainOP.calc_loss / calc_loss_rollout run on
    the reference's own class: loss, [loss, rec, kld], every parameter's gradient norm."""
    from egogen_b200.assets import fill_params_
    from oracle import cvae_train as oc, nets
    g = np.load(os.path.join(golden_dir, "train_golden.npz"))
    t = lambda k: torch.as_tensor(g[k])
    pred = nets.PredictorOracle().train()
    fill_params_(pred, seed=41)
    loss, rec, kld, _ = oc.primitive_loss(pred, t("p_data")[:2], t("p_data")[2:], t("p_eps"))
    loss.backward()
    assert _close([loss.item(), rec.item(), kld.item()], g["p_info"], 1e-5) and _close(loss.item(), g["p_loss"], 1e-5)
    assert _close(_gradnorms(pred), g["p_gradnorm"], 2e-4)
    assert np.allclose(pred.d_out.bias.grad.numpy(), g["p_grad_dout_bias"], rtol=1e-4, atol=1e-7)
    pred.zero_grad()
    loss = oc.rollout_loss(pred, t("r_markers"), t("r_jts"), list(t("r_eps")))
    loss.backward()
    assert _close(loss.item(), g["r_loss"], 1e-5) and _close(loss.item(), g["r_info"][0], 1e-5)
    assert _close(_gradnorms(pred), g["r_gradnorm"], 5e-4)


def test_regressor_training_oracle_matches_reference_trainop(golden_dir, smplx_model):
    from egogen_b200 import assets
    from oracle import cvae_train as oc, nets
    from oracle.smplx_lbs import SMPLXParserOracle
    g = np.load(os.path.join(golden_dir, "train_golden.npz"))
    reg = nets.RegressorOracle().train()
    _regressor_weights(reg, 3)
    lbs = SMPLXParserOracle(smplx_model, marker=assets.marker_ids())
    xb, loss, lm, lh = oc.regressor_loss(reg, lbs, torch.as_tensor(g["g_marker_ref"]), torch.as_tensor(g["g_betas"]))
    loss.backward()
    assert np.allclose(xb.detach().numpy(), g["g_xb"], rtol=1e-5, atol=1e-6)
    assert _close(loss.item(), g["g_loss"], 1e-5) and _close([lm.item(), lh.item()], g["g_items"], 1e-5)
    assert _close(_gradnorms(reg), g["g_gradnorm"], 2e-4)
    assert np.allclose(reg.pnet.out_fc.bias.grad.numpy(), g["g_grad_out_bias"], rtol=1e-4, atol=1e-8)


def test_combo_oracle_matches_reference_trainop(golden_dir, smplx_model):
    from egogen_b200 import assets
    from egogen_b200.assets import fill_params_
    from oracle import cvae_train as oc, nets
    from oracle.smplx_lbs import SMPLXParserOracle
    g = np.load(os.path.join(golden_dir, "train_golden.npz"))
    lbs = SMPLXParserOracle(smplx_model, marker=assets.marker_ids())
    data, betas, eps = torch.as_tensor(g["p_data"]), torch.as_tensor(g["c_betas"]), torch.as_tensor(g["c_eps"])
    for sched in (0, 1):
        pred, reg = nets.PredictorOracle().train(), nets.RegressorOracle().train()
        fill_params_(pred, seed=41)
        _regressor_weights(reg, 5)
        loss, items, _ = oc.combo_loss_one(pred, reg, lbs, data[:2], data[2:], betas[2:], eps, scheduled_sampling=bool(sched))
        loss.backward()
        assert _close(loss.item(), g[f"c{sched}_loss"], 1e-5), sched
        assert _close([x.item() for x in items], g[f"c{sched}_info"], 1e-5), sched
        assert _close(_gradnorms(pred), g[f"c{sched}_gradnorm"], 5e-4), sched
        assert int(g[f"c{sched}_reg_has_grad"]) == 1        # autograd reaches the regressor too; only the predictor is stepped


def test_env_oracle_matches_reference_crowd_env(golden_dir):
    """oracle.env.CrowdEnvOracle (batched, dup removed) vs trajectories of the reference's OWN CrowdEnv.reset / step run on
    the CPU (tests/golden/gen_env_golden.py): observations, rewards, termination and the carried state (body seed, R0, T0)
    over reset + 3 steps of 3 envs, pre-training and fine-tuning reward settings."""
    from egogen_b200 import assets
    from oracle import harness
    g = np.load(os.path.join(golden_dir, "env_golden.npz"))
    assert g["feet_vids_sorted"].tolist() == sorted(assets.feet_vids())
    world = harness.build_oracle_world(0, sdf_res=64)
    wp, goals, betas = harness.sample_candidates_cpu(world, 3, seed=5)
    assert np.array_equal(wp.numpy(), g["wp"]) and np.array_equal(goals.numpy(), g["goals"])
    Z = torch.as_tensor(g["Z"])
    env = world["env"]
    worst = {}

    def check(name, got, ref, tol):
        err = float(np.abs(np.asarray(got, dtype=np.float64) - np.asarray(ref, dtype=np.float64)).max())
        worst[name] = max(worst.get(name, 0.0), err)
        assert err <= tol, (name, err)

    for fin in (0, 1):
        env.finetuning = bool(fin)
        r = env.reset_from(wp, goals, betas)
        assert bool(r["accept"].all())
        env.set_state(state=r["state"], seed=r["seed"], R0=r["R0"], T0=r["T0"], betas=betas, dist=r["dist"],
                      steps=torch.zeros(3, dtype=torch.int64), goal=goals)
        for e in range(3):
            k = f"f{fin}_e{e}_"
            check("state0", r["state"][e], g[k + "state"][0], 2e-5)
            check("ego0", r["egosensing"][e], g[k + "ego"][0], 1e-5)
            check("dist0", r["obs_dist"][e], g[k + "dist"][0, 0], 1e-6)
            check("seed0", r["seed"][e], g[k + "seed"][0], 2e-5)
            check("R0_0", r["R0"][e], g[k + "R0"][0], 1e-6)
            check("T0_0", r["T0"][e], g[k + "T0"][0], 1e-6)
            assert float(g[k + "time"][0, 0]) == 1.0
        for s in range(3):
            o = env.step(Z[:, s])
            for e in range(3):
                k = f"f{fin}_e{e}_"
                check("state", o["state"][e], g[k + "state"][s + 1], 1e-4)
                check("ego", o["egosensing"][e], g[k + "ego"][s + 1], 1e-3)
                check("dist", o["dist"][e], g[k + "dist"][s + 1, 0], 1e-5)
                check("time", o["time"][e], g[k + "time"][s + 1, 0], 1e-6)
                check("reward", o["reward"][e], g[k + "reward"][s], 1e-4)
                check("seed", o["seed"][e], g[k + "seed"][s + 1], 1e-4)
                check("R0", o["R0"][e], g[k + "R0"][s + 1], 1e-5)
                check("T0", o["T0"][e], g[k + "T0"][s + 1], 1e-4)
                assert bool(o["terminated"][e]) == bool(g[k + "term"][s]), (fin, e, s)
    print("max |oracle - reference CrowdEnv|:", {k: f"{v:.2e}" for k, v in worst.items()})


def test_ppo_oracle_matches_reference_policy_learn(golden_dir):
    """oracle.ppo.learn_minibatch + clip_and_adamw vs the reference's own GAMMAPPOPolicy.forward / learn (one minibatch:
    clip / value / entropy terms, backward, gradient clip over actor+critic only, AdamW) run over a tianshou shim
    (tests/golden/gen_ppo_golden.py)."""
    from oracle import ppo as oppo
    g = np.load(os.path.join(golden_dir, "ppo_golden.npz"))
    actor = fill_params_(nets.ActorOracle(), seed=21)
    critic = fill_params_(nets.CriticOracle(), seed=22)
    shared = fill_params_(nets.PolicyBaseOracle(), seed=23)
    with torch.no_grad():
        for mod in actor.pnet.modules():
            if isinstance(mod, torch.nn.Linear):
                mod.weight.mul_(0.01)
    obs = {k: _t(g[k]) for k in ("state", "egosensing", "dist", "time")}
    optim = torch.optim.AdamW(list(actor.parameters()) + list(critic.parameters()) + list(shared.parameters()), lr=3e-4,
                              weight_decay=0.01)
    r = oppo.learn_minibatch(actor, critic, shared, obs, _t(g["act"]), _t(g["logp_old"]), _t(g["adv"]), _t(g["returns"]),
                             eps_clip=0.1, vf_coef=1.0, ent_coef=0.01, norm_adv=True)
    assert torch.allclose(r["mu"], _t(g["z_mu"]), atol=1e-6, rtol=1e-5) and torch.allclose(r["logvar"], _t(g["z_logvar"]), atol=1e-6, rtol=1e-5)
    assert torch.allclose(r["logp"], _t(g["logp_fw"]), atol=1e-3, rtol=1e-6)
    for k in ("loss", "clip", "vf", "ent", "kld"):
        assert abs(r[k] - float(g[k])) <= 1e-5 * max(1.0, abs(float(g[k]))), (k, r[k], float(g[k]))
    oppo.clip_and_adamw(actor, critic, shared, max_grad_norm=0.1, optim=optim)
    gn = lambda m: np.array([p.grad.norm().item() for p in m.parameters()])     # as learn() leaves them: clipped in place
    for name, m in (("actor", actor), ("critic", critic), ("shared", shared)):
        assert np.allclose(gn(m), g[name + "_gradnorm"], rtol=5e-4, atol=1e-9), (name, gn(m), g[name + "_gradnorm"])
    pn = lambda m: np.array([p.detach().norm().item() for p in m.parameters()])
    for name, m in (("actor", actor), ("critic", critic), ("shared", shared)):
        assert np.allclose(pn(m), g[name + "_after"], rtol=1e-6, atol=1e-9), name
    assert np.allclose(actor.pnet.out_fc.weight.detach()[:8, :16].numpy(), g["actor_out_w_after"], rtol=1e-5, atol=1e-8)


def test_box_env_oracle_matches_reference_box_env(golden_dir):
    """box_mode of the oracle env vs the reference's own crowd_env_2f_box.CrowdEnv (get_map walkability grid from the
    navmesh triangles, marker-bbox penetration count, termination on penetration) on the same starts and actions."""
    from egogen_b200 import assets
    from oracle import harness
    from oracle.env import CrowdEnvOracle
    g = np.load(os.path.join(golden_dir, "env_golden.npz"))
    world = harness.build_oracle_world(0, sdf_res=64)
    base = world["env"]
    tris = assets.scene_navmesh_triangles(assets.make_box_scene(0))
    env = CrowdEnvOracle(base.parser, base.combo, base.vposer, base.sdf, base.segments, base.marker, base.feet_marker_idx,
                         base.feet_vids, box_mode=True, navmesh_tris=tris, weight_look=0.1)
    wp, goals, betas, Z = (torch.as_tensor(g[k]) for k in ("wp", "goals", "betas", "Z"))
    r = env.reset_from(wp, goals, betas)
    assert bool(r["accept"].all())
    env.set_state(state=r["state"], seed=r["seed"], R0=r["R0"], T0=r["T0"], betas=betas, dist=r["dist"],
                  steps=torch.zeros(3, dtype=torch.int64), goal=goals)
    err = lambda a, b: float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max())
    done = [False] * 3
    for e in range(3):
        assert err(r["state"][e], g[f"box_e{e}_state"][0]) < 2e-5 and err(r["egosensing"][e], g[f"box_e{e}_ego"][0]) < 1e-5
    n_cmp = 0
    for s in range(3):
        o = env.step(Z[:, s])
        for e in range(3):
            k = f"box_e{e}_"
            if done[e] or s >= len(g[k + "reward"]):
                continue
            assert err(o["state"][e], g[k + "state"][s + 1]) < 1e-4, (e, s)
            assert err(o["egosensing"][e], g[k + "ego"][s + 1]) < 1e-3, (e, s)
            assert err(o["reward"][e], g[k + "reward"][s]) < 1e-4, (e, s, float(o["reward"][e]), g[k + "reward"][s])
            assert err(o["seed"][e], g[k + "seed"][s + 1]) < 1e-4 and err(o["T0"][e], g[k + "T0"][s + 1]) < 1e-4
            assert bool(o["terminated"][e]) == bool(g[k + "term"][s]), (e, s)
            done[e] = bool(g[k + "term"][s])
            n_cmp += 1
    assert n_cmp >= 3


def test_crowd_env_oracle_matches_reference_crowd_env(golden_dir):
    """4-agent crowd dynamics vs the reference's own crowd_env_crowd_eval.CrowdEnv driven in DummyCrowdVectorEnv's order:
    the other agents' marker boxes are holes of the floor polygon (walkability map + ego rays), redistributed before each
    agent's step; no penetration termination."""
    from egogen_b200 import assets
    from oracle import harness
    from oracle.env import CrowdEnvOracle
    g = np.load(os.path.join(golden_dir, "env_golden.npz"))
    world = harness.build_oracle_world(0, sdf_res=64)
    base = world["env"]
    A = 4
    fl = np.array([[4, 4], [4, -4], [-4, -4], [-4, 4], [4, 4]], np.float64)          # crowd_env_crowd_eval.py:391
    tris = np.stack([fl[[0, 1, 2]], fl[[2, 3, 0]]]).astype(np.float32)
    wp, goals, betas, Z = (torch.as_tensor(g[k]) for k in ("crowd_wp", "crowd_goals", "crowd_betas", "crowd_Z"))
    envs = []
    for a in range(A):
        e = CrowdEnvOracle(base.parser, base.combo, base.vposer, base.sdf, assets.rings_to_segments([fl]), base.marker,
                           base.feet_marker_idx, base.feet_vids, box_mode=True, navmesh_tris=tris, weight_look=0.1)
        e.crowd = True
        envs.append(e)
    sl = lambda a: slice(a, a + 1)
    bb = [envs[a].reset_from(wp[sl(a)], goals[sl(a)], betas[sl(a)])["bbox"] for a in range(A)]       # [1,4] each
    holes_for = lambda a: torch.stack([bb[o] for o in range(A) if o != a], dim=1)                      # [1,A-1,4]
    err = lambda x, y: float(np.abs(np.asarray(x, dtype=np.float64) - np.asarray(y, dtype=np.float64)).max())
    assert err(torch.cat(bb), g["crowd_bbox"][0]) < 1e-5
    for a in range(A):
        envs[a].holes = holes_for(a)
        r = envs[a].reset_from(wp[sl(a)], goals[sl(a)], betas[sl(a)])
        assert err(r["state"][0], g["crowd_state"][0, a]) < 2e-5 and err(r["egosensing"][0], g["crowd_ego"][0, a]) < 1e-4, a
        envs[a].set_state(state=r["state"], seed=r["seed"], R0=r["R0"], T0=r["T0"], betas=betas[sl(a)], dist=r["dist"],
                          steps=torch.zeros(1, dtype=torch.int64), goal=goals[sl(a)])
    envs[0].holes = None                                        # the other agents must matter in this set-up
    blind = envs[0].reset_from(wp[sl(0)], goals[sl(0)], betas[sl(0)])["egosensing"][0]
    assert err(blind, g["crowd_ego"][0, 0]) > 0.05
    for s in range(2):
        for a in range(A):
            envs[a].holes = holes_for(a)                        # update_holes_for_each_agent() before this worker's step
            o = envs[a].step(Z[s, a:a + 1])
            bb[a] = envs[a].bbox
            assert err(o["state"][0], g["crowd_state"][s + 1, a]) < 1e-4, (s, a)
            assert err(o["egosensing"][0], g["crowd_ego"][s + 1, a]) < 1e-3, (s, a)
            assert err(o["reward"][0], g["crowd_reward"][s, a]) < 1e-4, (s, a, float(o["reward"][0]), g["crowd_reward"][s, a])
            assert err(envs[a].bbox[0], g["crowd_bbox"][s + 1, a]) < 1e-4 and err(o["seed"][0], g["crowd_seed"][s, a]) < 1e-4
            assert bool(o["terminated"][0]) == bool(g["crowd_term"][s, a])


def test_start_body_oracle_matches_reference_sampler(golden_dir, smplx_model):
    """oracle.sampler.gen_init_body vs the reference's own CrowdMotion.gen_init_body (environments.py:1041-1131) run on
    its motion seed subseq_00343 (tests/golden/gen_sampler_golden.py): face-the-goal rotation, yaw jitter, pelvis-over-
    start / feet-on-floor translation, way-point heights; the motion-seed fixture carries the same frames."""
    from scipy.spatial.transform import Rotation
    from egogen_b200 import assets
    from oracle import sampler as osampler
    g = np.load(os.path.join(golden_dir, "sampler_golden.npz"))
    seed = np.load(os.path.join(golden_dir, "locomotion_seed_00343.npz"))
    parser = SMPLXParserOracle(smplx_model, marker=assets.marker_ids())
    for i in range(3):
        sf = int(g["start_frame"][i])
        assert np.array_equal(seed["poses"][sf:sf + 2, 3:66], g["body_pose"][i]) and np.array_equal(seed["betas"][:10], g["betas"][i])
        out = osampler.gen_init_body(parser, g["start"][i], g["target"][i], seed["betas"][:10], seed["poses"][sf:sf + 2, 3:66],
                                     seed["poses"][sf:sf + 2, :3], seed["trans"][sf:sf + 2], float(g["yaw"][i]))
        assert np.abs(out["transl"].numpy() - g["transl"][i]).max() < 2e-5, i
        assert np.abs(out["wpath"].numpy() - g["wpath"][i]).max() < 2e-5, i
        R_ref = Rotation.from_rotvec(g["global_orient"][i].astype(np.float64)).as_matrix()
        assert np.abs(out["global_orient_matrix"].numpy() - R_ref).max() < 2e-5, i
        # the body faces its goal up to the yaw jitter, pelvis above the start point, lowest joint on the floor
        assert np.abs(out["joints"][0, 0, :2].numpy() - g["start"][i][:2]).max() < 1e-4
        assert abs(float(out["joints"][0, :, 2].min())) < 1e-4


def test_env_oracle_penetration_termination_matches_reference(golden_dir):
    """A start that walks into the box: in the fine-tuning setting the reference terminates the episode on the per-frame
    penetration count (crowd_env_2f.py:175-176,299-302); the oracle must end the same step with the same reward."""
    from oracle import harness
    g = np.load(os.path.join(golden_dir, "env_golden.npz"))
    world = harness.build_oracle_world(0, sdf_res=64)
    env = world["env"]
    env.finetuning = True
    wp, goal, betas = torch.as_tensor(g["pen_wp"])[None], torch.as_tensor(g["pen_goal"])[None], torch.as_tensor(g["pen_betas"])[None]
    r = env.reset_from(wp, goal, betas)
    assert bool(r["accept"][0]) and np.abs(r["state"][0].numpy() - g["pen_state"][0]).max() < 2e-5
    env.set_state(state=r["state"], seed=r["seed"], R0=r["R0"], T0=r["T0"], betas=betas, dist=r["dist"],
                  steps=torch.zeros(1, dtype=torch.int64), goal=goal)
    Z = torch.as_tensor(g["Z"])
    n = len(g["pen_term"])
    assert bool(g["pen_term"][-1]) and n < 13
    for s in range(n):
        o = env.step(Z[0:1, s])
        assert np.abs(o["state"][0].numpy() - g["pen_state"][s + 1]).max() < 1e-4
        assert abs(float(o["reward"][0]) - float(g["pen_reward"][s])) < 1e-4
        assert bool(o["terminated"][0]) == bool(g["pen_term"][s])
    assert int(o["counts"][0].max()) >= 40 and float(o["dist"][0]) < 1.0       # ended by penetration, not by the goal / depth
    env.finetuning = False


def test_box_env_oracle_penetration_termination_matches_reference(golden_dir):
    """box-scene env: a start the 2-D map test accepts that ends by the marker-bbox penetration count
    (crowd_env_2f_box.py:279-295,325); same step, same reward (r_pene drops to 0)."""
    from egogen_b200 import assets
    from oracle import harness
    from oracle.env import CrowdEnvOracle
    g = np.load(os.path.join(golden_dir, "env_golden.npz"))
    assert "penbox_wp" in g.files
    world = harness.build_oracle_world(0, sdf_res=64)
    base = world["env"]
    tris = assets.scene_navmesh_triangles(assets.make_box_scene(0))
    env = CrowdEnvOracle(base.parser, base.combo, base.vposer, base.sdf, base.segments, base.marker, base.feet_marker_idx,
                         base.feet_vids, box_mode=True, navmesh_tris=tris, weight_look=0.1)
    wp, goal, betas = torch.as_tensor(g["penbox_wp"])[None], torch.as_tensor(g["penbox_goal"])[None], torch.as_tensor(g["pen_betas"])[None]
    r = env.reset_from(wp, goal, betas)
    assert bool(r["accept"][0]) and np.abs(r["state"][0].numpy() - g["penbox_state"][0]).max() < 2e-5
    env.set_state(state=r["state"], seed=r["seed"], R0=r["R0"], T0=r["T0"], betas=betas, dist=r["dist"],
                  steps=torch.zeros(1, dtype=torch.int64), goal=goal)
    Z = torch.as_tensor(g["Z"])
    n = len(g["penbox_term"])
    assert bool(g["penbox_term"][-1]) and n < 13
    for s in range(n):
        o = env.step(Z[0:1, s])
        assert np.abs(o["state"][0].numpy() - g["penbox_state"][s + 1]).max() < 1e-4
        assert abs(float(o["reward"][0]) - float(g["penbox_reward"][s])) < 1e-4
        assert bool(o["terminated"][0]) == bool(g["penbox_term"][s])
    assert float(o["terms"][0, 6]) == 0.0 and float(o["dist"][0]) < 1.0        # r_pene = 0: ended by penetration


def test_canonicalisation_oracle_matches_reference_script(golden_dir, smplx_model):
    """oracle.sampler.canonicalize_subsequence vs the primitive the reference's own canonicalize_subsequence produced for
    the same synthetic recording (tests/golden/gen_canon_golden.py); the GPU test of egogen_b200.primitive_batches compares
    the CUDA path with this same pipeline on this same recording."""
    from scipy.spatial.transform import Rotation
    from egogen_b200 import assets
    from oracle import sampler as osampler
    g = np.load(os.path.join(golden_dir, "canon_golden.npz"))
    p67 = SMPLXParserOracle(smplx_model, marker=assets.marker_ids())
    p41 = SMPLXParserOracle(smplx_model, marker=assets.marker_ids("cmu_41"))
    assert osampler.canonicalize_subsequence(p67, p41, g["in_betas"], g["in_transl"], g["in_pose"], 150, 210) is None
    out = osampler.canonicalize_subsequence(p67, p41, g["in_betas"], g["in_transl"], g["in_pose"], 30, 90)
    assert np.abs(out["transf_rotmat"] - g["transf_rotmat"]).max() < 1e-6 and np.abs(out["transf_transl"] - g["transf_transl"]).max() < 1e-6
    assert np.abs(out["trans"] - g["trans"]).max() < 2e-6
    rot = lambda aa: Rotation.from_rotvec(np.asarray(aa, dtype=np.float64)).as_matrix()
    assert np.abs(rot(out["poses"][:, :3]) - rot(g["poses"][:, :3])).max() < 2e-6       # same rotation (tgm vs scipy axis-angle)
    assert np.array_equal(out["poses"][:, 3:], g["poses"][:, 3:].astype(np.float32))
    for k in ("joints", "marker_ssm2_67", "marker_cmu_41"):
        assert out[k].shape == g[k].shape and np.abs(out[k] - g[k]).max() < 5e-6, k


def test_train_loop_lr_rule_matches_reference_scheduler(golden_dir):
    """lr_at(epoch) of the train-op mirrors vs the reference's get_scheduler('lambda') stepped once per epoch."""
    from egogen_b200.train_gamma_predictor import GAMMAPrimitiveVAETrainOP
    from egogen_b200.train_gamma_regressor import GAMMARegressorTrainOP
    g = np.load(os.path.join(golden_dir, "train_golden.npz"))
    cfg = {"learning_rate": 5e-4, "num_epochs": 400, "num_epochs_fix": 100}
    for cls in (GAMMAPrimitiveVAETrainOP, GAMMARegressorTrainOP):
        op = cls(trainconfig=cfg, device="cpu")
        mine = np.array([op.lr_at(e) for e in range(400)])
        assert np.allclose(mine, g["sched_lr"], rtol=1e-12, atol=0), cls.__name__
    assert g["sched_lr"][0] == 5e-4 and g["sched_lr"][100] == 5e-4 and g["sched_lr"][399] < 5e-6
